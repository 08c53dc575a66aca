// K2: query x corpus score kernel for sm_100a.
//
//   S[q, v] = sum_k A[q, k] * B[v, k]        A = query operand, B = corpus operand (bf16, K-major)
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B swizzle, 4-stage smem ring)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1, 128 x 256 x 16, fp32 accumulators in TMEM)
//   warp 2      TMEM allocator (512 columns = two 128 x 256 accumulator stages)
//   warps 4-7   epilogue       (tcgen05.ld, one query row per thread)
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile
// i+1.  Two epilogues share the main loop:
//   STORE   out = alpha * S                                   (cal_error / sampling pass)
//   FILTER  per-row window (lo, hi]: count scores above hi, append (score, index) of scores inside
//           the window to a per-row candidate list -- the score matrix never reaches HBM.
// Work is cut into units (one 128-query tile x a run of corpus tiles), ordered corpus-chunk-major so
// the CTAs that run concurrently stream the same corpus rows and share them through L2, while the
// query operand (tens of MB) stays L2-resident.
//
// Replaces np.dot(l2norm(captions), l2norm(videos).T) in LINAS-engine/evaluation.py:21,45,79 and
// P @ index.T in MultiFusion/src/validate.py:73,90, and (FILTER) the per-row argsort that follows
// them (LINAS-engine/inference.py:79, MultiFusion/src/validate.py:74,92).
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace xmve {
namespace {

constexpr int BM = 128;              // query rows per tile (TMEM lanes)
constexpr int BN = 256;              // corpus rows per tile (TMEM columns per accumulator stage)
constexpr int BK = 64;               // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int TMEM_COLS = ACC_STAGES * BN;   // 512
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

enum { MODE_STORE = 0, MODE_FILTER = 1 };

struct Params {
  int64_t nq, nv;
  int k_blocks;
  int m_tiles, n_tiles, tiles_per_unit;
  int64_t n_units;
  // STORE
  float alpha;
  float* out;
  int64_t out_ld;
  int vec_ok;
  // FILTER
  const float* lo;
  const float* hi;
  int32_t* count_above;
  int32_t* cand_count;
  float* cand_score;
  int32_t* cand_idx;
  int32_t cap;
};

struct Unit {
  int mt, t0, t1;
};
__device__ __forceinline__ Unit decode_unit(const Params& p, int64_t u) {
  Unit x;
  const int chunk = static_cast<int>(u / p.m_tiles);
  x.mt = static_cast<int>(u - static_cast<int64_t>(chunk) * p.m_tiles);
  x.t0 = chunk * p.tiles_per_unit;
  x.t1 = min(p.n_tiles, x.t0 + p.tiles_per_unit);
  return x;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
score_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5;   // warp-uniform
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 4 * 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const Unit un = decode_unit(p, u);
        for (int t = un.t0; t < un.t1; ++t) {
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * STAGE_BYTES;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            ptx::tma_load_2d(&tm_a, &full_bar[stage], sa, kb * BK, un.mt * BM);
            ptx::tma_load_2d(&tm_b, &full_bar[stage], sa + A_BYTES, kb * BK, t * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(BM, BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const Unit un = decode_unit(p, u);
        for (int t = un.t0; t < un.t1; ++t) {
          ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);     // epilogue has drained this accumulator
          ptx::tc_fence_after_sync();
          const uint32_t tmem_d = tmem_base + acc * BN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);           // TMA bytes have landed
            ptx::tc_fence_after_sync();
            const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t da = ptx::smem_desc_k_sw128(sa);
            const uint64_t db = ptx::smem_desc_k_sw128(sa + A_BYTES);
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
              // +32 bytes along K inside the 128-byte swizzle row = +2 in the (addr >> 4) field
              ptx::umma_bf16(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
            }
            ptx::umma_commit(&empty_bar[stage]);               // frees the smem slot when the MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit(&tfull_bar[acc]);                   // accumulator complete -> epilogue
          if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ================================ epilogue ====================================
    const int quarter = warp & 3;                              // TMEM lane quarter this warp may read
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit un = decode_unit(p, u);
      const int64_t row = static_cast<int64_t>(un.mt) * BM + row_in_tile;
      const bool row_ok = row < p.nq;
      float lo = __int_as_float(0x7f800000), hi = __int_as_float(0x7f800000);
      int cnt = 0;
      if (MODE == MODE_FILTER && row_ok) {
        lo = p.lo[row];
        if (p.hi != nullptr) hi = p.hi[row];
      }
      for (int t = un.t0; t < un.t1; ++t) {
        ptx::mbar_wait(&tfull_bar[acc], acc_phase);
        ptx::tc_fence_after_sync();
        const int64_t col0 = static_cast<int64_t>(t) * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + lane_addr + acc * BN + c * 32, v);
          ptx::tmem_ld_wait();
          const int64_t col = col0 + c * 32;
          if (MODE == MODE_STORE) {
            if (row_ok) {
              float* dst = p.out + row * p.out_ld + col;
              if (p.vec_ok && col + 32 <= p.nv) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  float4 w = make_float4(p.alpha * __uint_as_float(v[i]), p.alpha * __uint_as_float(v[i + 1]),
                                         p.alpha * __uint_as_float(v[i + 2]), p.alpha * __uint_as_float(v[i + 3]));
                  *reinterpret_cast<float4*>(dst + i) = w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col + i < p.nv) dst[i] = p.alpha * __uint_as_float(v[i]);
              }
            }
          } else {
            float m = __uint_as_float(v[0]);
#pragma unroll
            for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
            if (m > lo) {                                      // rare: at least one score enters the window
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float s = __uint_as_float(v[i]);
                if (s > lo && col + i < p.nv) {
                  if (s > hi) {
                    ++cnt;
                  } else {
                    const int slot = atomicAdd(&p.cand_count[row], 1);
                    if (slot < p.cap) {
                      p.cand_score[row * p.cap + slot] = s;
                      p.cand_idx[row * p.cap + slot] = static_cast<int32_t>(col + i);
                    }
                  }
                }
              }
            }
          }
        }
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive(&tempty_bar[acc]);                    // 128 arrivals release the accumulator
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
      if (MODE == MODE_FILTER && cnt != 0 && p.count_above != nullptr) atomicAdd(&p.count_above[row], cnt);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (cached == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    XMVE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess)
      return fail(XMVE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cached = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = cached;
  return XMVE_OK;
}

// bf16 [rows, k] with row stride `ld` elements -> 2-D tiled map, box = 64 x box_rows, 128B swizzle,
// out-of-bounds elements read as zero.
int make_operand_map(CUtensorMap* map, const void* base, int64_t rows, int64_t k, int64_t ld, int box_rows) {
  EncodeTiledFn encode;
  int s = get_encode_fn(&encode);
  if (s != XMVE_OK) return s;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(XMVE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld k=%lld ld=%lld", static_cast<int>(r),
                static_cast<long long>(rows), static_cast<long long>(k), static_cast<long long>(ld));
  return XMVE_OK;
}

int check_operands(const void* a_op, int64_t nq, int64_t a_ld, const void* b_op, int64_t nv, int64_t b_ld,
                   int64_t b_row_step, int k) {
  XMVE_REQUIRE(a_op != nullptr && b_op != nullptr, "score: null operand");
  XMVE_REQUIRE(nq > 0 && nv > 0, "score: empty problem (nq=%lld nv=%lld)", (long long)nq, (long long)nv);
  XMVE_REQUIRE(k > 0 && k % BK == 0, "score: k=%d must be a positive multiple of %d", k, BK);
  XMVE_REQUIRE(a_ld >= k && b_ld >= k && a_ld % 8 == 0 && b_ld % 8 == 0,
               "score: row strides must be >= k and multiples of 8 elements");
  XMVE_REQUIRE(aligned16(a_op) && aligned16(b_op), "score: operands must be 16-byte aligned");
  XMVE_REQUIRE(b_row_step >= 1, "score: b_row_step must be >= 1");
  if (nv > (int64_t(1) << 31) - BN || nq > (int64_t(1) << 31) - BM)
    return fail(XMVE_ERR_LIMIT, "score: more than 2^31 rows");
  return XMVE_OK;
}

template <int MODE>
int launch(const void* a_op, int64_t nq, int64_t a_ld, const void* b_op, int64_t nv, int64_t b_ld, int64_t b_row_step,
           int k, Params p, cudaStream_t stream) {
  CUtensorMap tm_a, tm_b;
  int s = make_operand_map(&tm_a, a_op, nq, k, a_ld, BM);
  if (s != XMVE_OK) return s;
  s = make_operand_map(&tm_b, b_op, nv, k, b_ld * b_row_step, BN);
  if (s != XMVE_OK) return s;

  const int sms = sm_count();
  if (sms <= 0) return fail(XMVE_ERR_DEVICE, "score: cannot query the SM count");
  p.nq = nq;
  p.nv = nv;
  p.k_blocks = k / BK;
  p.m_tiles = static_cast<int>((nq + BM - 1) / BM);
  p.n_tiles = static_cast<int>((nv + BN - 1) / BN);
  // aim for >= 16 units per CTA (tail balance) with at most 64 corpus tiles per unit
  int64_t tpu = (static_cast<int64_t>(p.n_tiles) * p.m_tiles) / (static_cast<int64_t>(sms) * 16);
  if (tpu < 1) tpu = 1;
  if (tpu > 64) tpu = 64;
  p.tiles_per_unit = static_cast<int>(tpu);
  const int64_t n_chunks = (p.n_tiles + tpu - 1) / tpu;
  p.n_units = n_chunks * p.m_tiles;
  const int grid = static_cast<int>(p.n_units < sms ? p.n_units : sms);

  static bool attr_set[2] = {false, false};
  if (!attr_set[MODE]) {
    XMVE_CUDA(cudaFuncSetAttribute(score_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set[MODE] = true;
  }
  score_kernel<MODE><<<grid, THREADS, SMEM_BYTES, stream>>>(tm_a, tm_b, p);
  return launch_status("score_kernel");
}

}  // namespace
}  // namespace xmve

extern "C" int xmve_score_store(const void* a_op, int64_t nq, int64_t a_ld, const void* b_op, int64_t nv,
                                int64_t b_ld, int64_t b_row_step, int k, float alpha, float* out, int64_t out_ld,
                                void* stream) {
  using namespace xmve;
  XMVE_DEVICE_OR_RETURN();
  int s = check_operands(a_op, nq, a_ld, b_op, nv, b_ld, b_row_step, k);
  if (s != XMVE_OK) return s;
  XMVE_REQUIRE(out != nullptr && out_ld >= nv, "score_store: out is null or out_ld < nv");
  Params p{};
  p.alpha = alpha;
  p.out = out;
  p.out_ld = out_ld;
  p.vec_ok = (out_ld % 4 == 0 && aligned16(out)) ? 1 : 0;
  return launch<MODE_STORE>(a_op, nq, a_ld, b_op, nv, b_ld, b_row_step, k, p, static_cast<cudaStream_t>(stream));
}

extern "C" int xmve_score_filter(const void* a_op, int64_t nq, int64_t a_ld, const void* b_op, int64_t nv,
                                 int64_t b_ld, int k, const float* lo, const float* hi, int32_t* count_above,
                                 int32_t* cand_count, float* cand_score, int32_t* cand_idx, int32_t cap,
                                 void* stream) {
  using namespace xmve;
  XMVE_DEVICE_OR_RETURN();
  int s = check_operands(a_op, nq, a_ld, b_op, nv, b_ld, 1, k);
  if (s != XMVE_OK) return s;
  XMVE_REQUIRE(lo != nullptr && cand_count != nullptr && cand_score != nullptr && cand_idx != nullptr && cap > 0,
               "score_filter: lo / candidate buffers are required and cap must be > 0");
  XMVE_REQUIRE(hi == nullptr || count_above != nullptr, "score_filter: hi given without count_above");
  Params p{};
  p.lo = lo;
  p.hi = hi;
  p.count_above = count_above;
  p.cand_count = cand_count;
  p.cand_score = cand_score;
  p.cand_idx = cand_idx;
  p.cap = cap;
  return launch<MODE_FILTER>(a_op, nq, a_ld, b_op, nv, b_ld, 1, k, p, static_cast<cudaStream_t>(stream));
}
