// Order statistics and final top-k selection over per-row candidate lists.
//
// These replace the full per-row sorts of the reference -- np.argsort(errors[0])[:topK]
// (LINAS-engine/inference.py:79) and torch.argsort(tmp.cpu(), dim=-1) (MultiFusion/src/validate.py:74,92)
// -- with (a) a radix select that turns a score sample / candidate list into a threshold and
// (b) a shared-memory bitonic sort of the few hundred survivors.  K3 (the G-way merge after the
// multi-GPU all-gather) is the same sort over int64 global indices.
#include <math_constants.h>

#include "common.cuh"

namespace xmve {
namespace {

__device__ __forceinline__ uint32_t float_key(float x) {      // ascending key order == ascending float order
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// j-th largest (1-based) of get(0..n); -inf if n < j.  All threads of the block must call it.
template <typename Get>
__device__ float block_kth_largest_of(Get get, int64_t n, int j, uint32_t* hist, uint32_t* bcast) {
  if (j <= 0 || n < j) return -CUDART_INF_F;
  uint32_t prefix = 0, mask = 0;
  uint32_t remaining = static_cast<uint32_t>(j);
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t k = float_key(get(i));
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t cum = 0;
      int b = 255;
      for (; b > 0; --b) {
        if (cum + hist[b] >= remaining) break;
        cum += hist[b];
      }
      bcast[0] = static_cast<uint32_t>(b);
      bcast[1] = remaining - cum;
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask |= 0xffu << shift;
    remaining = bcast[1];
    __syncthreads();
  }
  return key_float(prefix);
}

__device__ float block_kth_largest(const float* __restrict__ vals, int64_t n, int j, uint32_t* hist, uint32_t* bcast) {
  return block_kth_largest_of([vals](int64_t i) { return vals[i]; }, n, j, hist, bcast);
}

__global__ void __launch_bounds__(256)
row_kth_kernel(const float* __restrict__ vals, int64_t cols, int64_t ld, const int32_t* __restrict__ counts,
               int j1, float sub, int j2, float* __restrict__ out) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bcast[2];
  const int64_t r = blockIdx.x;
  int64_t n = cols;
  if (counts != nullptr) n = min(static_cast<int64_t>(counts[r]), cols);
  const float* row = vals + r * ld;
  float res = block_kth_largest(row, n, j1, hist, bcast) - sub;
  if (j2 > 0) res = fmaxf(res, block_kth_largest(row, n, j2, hist, bcast));
  if (threadIdx.x == 0) out[r] = res;
}

// The J largest values of each row, descending, -inf padded.  Used to combine order statistics across corpus
// shards: the j-th largest of a union is among the per-shard top-j lists.
__global__ void __launch_bounds__(256)
row_topj_kernel(const float* __restrict__ vals, int64_t cols, int64_t ld, const int32_t* __restrict__ counts, int J,
                int P, float* __restrict__ out) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bcast[2];
  __shared__ int n_buf;
  extern __shared__ float topj_buf[];
  const int64_t r = blockIdx.x;
  int64_t n = cols;
  if (counts != nullptr) n = min(static_cast<int64_t>(counts[r]), cols);
  const float* row = vals + r * ld;
  if (threadIdx.x == 0) n_buf = 0;
  const float t = block_kth_largest(row, n, J, hist, bcast);       // -inf when the row has fewer than J values
  __syncthreads();
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = row[i];
    if (v > t) topj_buf[atomicAdd(&n_buf, 1)] = v;                  // strictly above the J-th largest: < J values
  }
  __syncthreads();
  const int c = n_buf;
  const int have = static_cast<int>(min(static_cast<int64_t>(J), n));
  for (int i = c + threadIdx.x; i < P; i += blockDim.x) topj_buf[i] = (i < have) ? t : -CUDART_INF_F;
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int x = threadIdx.x; x < P; x += blockDim.x) {
        const int o = x ^ stride;
        if (o > x) {
          const bool desc = (x & size) == 0;
          const float a = topj_buf[x], b = topj_buf[o];
          if (desc ? (b > a) : (a > b)) {
            topj_buf[x] = b;
            topj_buf[o] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < J; i += blockDim.x) out[r * J + i] = topj_buf[i];
}

// ---- bitonic top-k ------------------------------------------------------------------------------
template <typename IdxT>
__device__ __forceinline__ bool before(double sa, IdxT ia, double sb, IdxT ib) {
  return sa > sb || (sa == sb && ia < ib);                    // score descending, index ascending
}

template <typename IdxT>
struct IdxLimits;
template <>
struct IdxLimits<int32_t> {
  static __device__ int32_t max() { return 0x7fffffff; }
};
template <>
struct IdxLimits<int64_t> {
  static __device__ int64_t max() { return 0x7fffffffffffffffLL; }
};

template <typename IdxT>
__global__ void __launch_bounds__(512)
select_topk_kernel(const double* __restrict__ score, const IdxT* __restrict__ idx, int64_t cols,
                   const int32_t* __restrict__ counts, int64_t idx_offset, const int64_t* __restrict__ exclude, int k,
                   const float* __restrict__ thr, float eps, const float* __restrict__ bound,
                   const int32_t* __restrict__ overflow, int pmax, double* __restrict__ out_score, int64_t* __restrict__ out_idx, int32_t* __restrict__ out_valid,
                   int32_t* __restrict__ cert, float* __restrict__ thr_next) {
  extern __shared__ __align__(16) uint8_t sel_smem[];
  double* keys = reinterpret_cast<double*>(sel_smem);
  IdxT* ids = reinterpret_cast<IdxT*>(keys + pmax);
  __shared__ int n_valid_s;
  const int64_t r = blockIdx.x;
  int64_t n_in = cols;
  bool cand_overflow = false;
  if (counts != nullptr) {
    cand_overflow = counts[r] > cols;
    n_in = min(static_cast<int64_t>(counts[r]), cols);
  }
  if (overflow != nullptr && overflow[r] != 0) cand_overflow = true;   // a shard's candidate list overflowed
  __shared__ uint32_t sel_hist[256];
  __shared__ uint32_t sel_bcast[2];
  if (threadIdx.x == 0) n_valid_s = 0;
  __syncthreads();
  const int64_t excl = exclude ? exclude[r] : -1;
  const double* srow = score + r * cols;
  const IdxT* irow = idx ? idx + r * cols : nullptr;
  // valid entry i -> its score, anything else -> -inf
  auto value = [&](int64_t i) -> double {
    const double s = srow[i];
    const IdxT id = irow ? irow[i] : static_cast<IdxT>(i);            // idx == NULL: column number
    return (s > -CUDART_INF && !(exclude && static_cast<int64_t>(id) + idx_offset == excl)) ? s : -CUDART_INF;
  };
  for (int64_t i = threadIdx.x; i < n_in; i += blockDim.x)
    if (value(i) > -CUDART_INF) atomicAdd(&n_valid_s, 1);
  __syncthreads();
  const int n_all = n_valid_s;                                         // every valid entry of the row
  // More valid entries than the sort holds (dense neighbourhoods: thousands of items within eps of the k-th
  // best): keep only those whose score, rounded to float, reaches the k-th largest rounded score.  Rounding is
  // monotone, so the kept set contains the exact top-k; it can exceed the capacity only through > pmax scores
  // that agree to float precision with the k-th.
  float floor_f = -CUDART_INF_F;
  if (n_all > pmax)
    floor_f = block_kth_largest_of([&](int64_t i) { return __double2float_rn(value(i)); }, n_in, k, sel_hist, sel_bcast);
  __syncthreads();
  if (threadIdx.x == 0) n_valid_s = 0;
  __syncthreads();
  for (int64_t i = threadIdx.x; i < n_in; i += blockDim.x) {
    const double s = value(i);
    if (s > -CUDART_INF && __double2float_rn(s) >= floor_f) {
      const int pos = atomicAdd(&n_valid_s, 1);
      if (pos < pmax) {
        keys[pos] = s;
        ids[pos] = irow ? irow[i] : static_cast<IdxT>(i);
      }
    }
  }
  __syncthreads();
  const int n_valid = n_valid_s;                                       // entries kept for the sort
  const int n = min(n_valid, pmax);
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = n + threadIdx.x; i < P; i += blockDim.x) {
    keys[i] = -CUDART_INF;
    ids[i] = IdxLimits<IdxT>::max();
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < P; t += blockDim.x) {
        const int o = t ^ stride;
        if (o > t) {
          const bool asc = (t & size) == 0;                   // "before" order towards lower indices
          const double sa = keys[t], sb = keys[o];
          const IdxT ia = ids[t], ib = ids[o];
          const bool swap = asc ? before<IdxT>(sb, ib, sa, ia) : before<IdxT>(sa, ia, sb, ib);
          if (swap) {
            keys[t] = sb; keys[o] = sa;
            ids[t] = ib; ids[o] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const bool ok = i < n;
    out_score[r * k + i] = ok ? keys[i] : -CUDART_INF;
    out_idx[r * k + i] = ok ? static_cast<int64_t>(ids[i]) + idx_offset : -1;
  }
  if (threadIdx.x == 0) {
    if (out_valid) out_valid[r] = n_all;
    if (cert != nullptr) {
      const float t = thr[r];
      const bool enough = n >= k && n_valid <= pmax;
      const double kth = enough ? keys[k - 1] : -CUDART_INF;
      const bool complete = (t == -CUDART_INF_F) || (kth - static_cast<double>(eps) >= static_cast<double>(t));
      cert[r] = (!cand_overflow && enough && complete) ? 1 : 0;
      if (thr_next != nullptr) {
        float nx;
        if (cand_overflow || n_valid > pmax) nx = bound ? bound[r] : t;          // approx kth of retained - 2 eps
        else if (enough) nx = static_cast<float>(kth) - 1.0001f * eps - 1e-7f;   // provably complete
        else nx = t - fmaxf(8.f * eps, 0.25f * fabsf(t));                        // too few found: lower and retry
        thr_next[r] = nx;
      }
    }
  }
}

int pow2_ceil(int64_t x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

template <typename IdxT>
int launch_select(const double* score, const IdxT* idx, int64_t rows, int64_t cols, const int32_t* counts,
                  int64_t idx_offset, const int64_t* exclude, int32_t k, const float* thr, float eps,
                  const float* bound, const int32_t* overflow, double* out_score, int64_t* out_idx, int32_t* out_valid,
                  int32_t* cert, float* thr_next, cudaStream_t st) {
  XMVE_REQUIRE(score && out_score && out_idx && rows >= 0 && cols > 0 && k > 0, "select_topk: bad arguments");
  XMVE_REQUIRE(cert == nullptr || thr != nullptr, "select_topk: cert needs thr");
  if (rows == 0) return XMVE_OK;
  int pmax = pow2_ceil(cols);
  if (pmax > 16384) pmax = 16384;
  if (pmax < k) return fail(XMVE_ERR_LIMIT, "select_topk: k=%d exceeds the %d-entry sort capacity", k, pmax);
  const int smem = pmax * static_cast<int>(sizeof(double) + sizeof(IdxT));
  if (smem > 220 * 1024) {
    pmax = 8192;                                            // int64 indices: 16 B per entry
    if (pmax < k) return fail(XMVE_ERR_LIMIT, "select_topk: k too large");
  }
  const int smem_bytes = pmax * static_cast<int>(sizeof(double) + sizeof(IdxT));
  static int attr_bytes[MAX_DEVICES] = {};                 // per <IdxT> instantiation and per device
  const int dev = current_device();
  if (dev < 0) return fail(XMVE_ERR_DEVICE, "select_topk: no current device");
  if (smem_bytes > attr_bytes[dev]) {
    XMVE_CUDA(cudaFuncSetAttribute(select_topk_kernel<IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_bytes[dev] = smem_bytes;
  }
  select_topk_kernel<IdxT><<<static_cast<unsigned>(rows), 512, smem_bytes, st>>>(
      score, idx, cols, counts, idx_offset, exclude, k, thr, eps, bound, overflow, pmax, out_score, out_idx, out_valid,
      cert, thr_next);
  return launch_status("select_topk_kernel");
}

}  // namespace
}  // namespace xmve

extern "C" int xmve_row_kth(const float* vals, int64_t rows, int64_t cols, int64_t ld, const int32_t* counts,
                            int32_t j1, float sub, int32_t j2, float* out, void* stream) {
  using namespace xmve;
  XMVE_DEVICE_OR_RETURN();
  XMVE_REQUIRE(vals && out && rows >= 0 && cols > 0 && ld >= cols && j1 > 0, "row_kth: bad arguments");
  if (rows == 0) return XMVE_OK;
  row_kth_kernel<<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(vals, cols, ld, counts,
                                                                                            j1, sub, j2, out);
  return launch_status("row_kth_kernel");
}

extern "C" int xmve_select_topk_i32(const double* score, const int32_t* idx, int64_t rows, int64_t cols,
                                    const int32_t* counts, int64_t idx_offset, const int64_t* exclude, int32_t k,
                                    const float* thr, float eps, const float* bound, double* out_score,
                                    int64_t* out_idx, int32_t* out_valid, int32_t* cert, float* thr_next,
                                    void* stream) {
  using namespace xmve;
  XMVE_DEVICE_OR_RETURN();
  return launch_select<int32_t>(score, idx, rows, cols, counts, idx_offset, exclude, k, thr, eps, bound, nullptr,
                                out_score, out_idx, out_valid, cert, thr_next, static_cast<cudaStream_t>(stream));
}

extern "C" int xmve_select_topk_i64(const double* score, const int64_t* idx, int64_t rows, int64_t cols,
                                    const int64_t* exclude, int32_t k, const float* thr, float eps,
                                    const int32_t* overflow, double* out_score, int64_t* out_idx, int32_t* out_valid,
                                    int32_t* cert, float* thr_next, void* stream) {
  using namespace xmve;
  XMVE_DEVICE_OR_RETURN();
  return launch_select<int64_t>(score, idx, rows, cols, nullptr, 0, exclude, k, thr, eps, nullptr, overflow, out_score,
                                out_idx, out_valid, cert, thr_next, static_cast<cudaStream_t>(stream));
}

extern "C" int xmve_row_topj(const float* vals, int64_t rows, int64_t cols, int64_t ld, const int32_t* counts,
                             int32_t j, float* out, void* stream) {
  using namespace xmve;
  XMVE_DEVICE_OR_RETURN();
  XMVE_REQUIRE(vals && out && rows >= 0 && cols > 0 && ld >= cols && j > 0, "row_topj: bad arguments");
  if (j > 4096) return fail(XMVE_ERR_LIMIT, "row_topj: j=%d exceeds 4096", j);
  if (rows == 0) return XMVE_OK;
  const int P = pow2_ceil(j);
  row_topj_kernel<<<static_cast<unsigned>(rows), 256, P * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      vals, cols, ld, counts, j, P, out);
  return launch_status("row_topj_kernel");
}
