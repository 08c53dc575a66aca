timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -5 gpurun_out/t_all.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/b10m_g2.log 2>&1; tail -1 gpurun_out/b10m_g2.log | cut -c1-1500
