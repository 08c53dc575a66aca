// K1: row normalise + cast.  HBM-bound: one warp per embedding row, the row is read from HBM once
// (the second pass hits L1/L2), norms are accumulated in fp64.
//
// Replaces evaluation.l2norm (LINAS-engine/evaluation.py:10-14), which cal_error re-runs over the
// whole corpus on every call (:19-20), F.normalize(index).float() (MultiFusion/src/validate.py:55)
// and Combiner.time_process (mean over frames, MultiFusion/src/combiner.py:140-143).
#include <cuda_bf16.h>

#include "common.cuh"

namespace xmve {
namespace {

constexpr int WARPS = 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// (frame-pooled) raw value of element i of one row, as fp32
template <typename T>
__device__ __forceinline__ float pooled(const T* __restrict__ r, int i, int d, int frames) {
  if (frames == 1) return static_cast<float>(r[i]);
  float s = 0.f;
  for (int f = 0; f < frames; ++f) s += static_cast<float>(r[static_cast<int64_t>(f) * d + i]);
  return s / static_cast<float>(frames);
}

__device__ __forceinline__ void write_planes(__nv_bfloat16* __restrict__ op, int i, int dpad, int layout, float y) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(y);
  if (layout == XMVE_OP_X1) {
    op[i] = hi;
    return;
  }
  const __nv_bfloat16 lo = __float2bfloat16_rn(y - __bfloat162float(hi));
  op[i] = hi;
  if (layout == XMVE_OP_X3_QUERY) {          // [hi | hi | lo]
    op[dpad + i] = hi;
    op[2 * dpad + i] = lo;
  } else {                                   // [hi | lo | hi]
    op[dpad + i] = lo;
    op[2 * dpad + i] = hi;
  }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
prepare_rows_kernel(const T* __restrict__ src, int64_t n, int d, int frames, int64_t src_ld,
                    float* __restrict__ raw_out, int64_t raw_ld, int64_t raw_off, double* __restrict__ norm_out,
                    float* __restrict__ resid_out, __nv_bfloat16* __restrict__ op_out, int64_t op_ld, int64_t op_off, int layout, int dpad,
                    float weight, int norm_mode, int vec4) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  if (row >= n) return;
  const T* __restrict__ r = src + row * src_ld;
  float* __restrict__ raw = raw_out ? raw_out + row * raw_ld + raw_off : nullptr;

  double ss = 0.0;
  if (vec4) {                                // fp32, frames == 1, 16-byte aligned rows, d % 4 == 0
    const float4* r4 = reinterpret_cast<const float4*>(r);
    for (int i = lane; i < d / 4; i += 32) {
      const float4 x = r4[i];
      ss += static_cast<double>(x.x) * x.x + static_cast<double>(x.y) * x.y + static_cast<double>(x.z) * x.z +
            static_cast<double>(x.w) * x.w;
      if (raw) reinterpret_cast<float4*>(raw)[i] = x;
    }
  } else {
    for (int i = lane; i < d; i += 32) {
      const float x = pooled(r, i, d, frames);
      ss += static_cast<double>(x) * static_cast<double>(x);
      if (raw) raw[i] = x;
    }
  }
  ss = warp_sum(ss);
  const double nrm = sqrt(ss);
  if (lane == 0 && norm_out) norm_out[row] = nrm;
  if (op_out == nullptr) return;

  const double den = (norm_mode == XMVE_NORM_EPS) ? fmax(nrm, 1e-12) : nrm;
  __nv_bfloat16* __restrict__ op = op_out + row * op_ld + op_off;
  const int planes = (layout == XMVE_OP_X1) ? 1 : 3;
  double rs = 0.0;                           // || w * x_hat - bf16(w * x_hat) ||^2 of this row (x1 operand)
  for (int i = lane; i < d; i += 32) {
    const float x = pooled(r, i, d, frames);
    const double xd = static_cast<double>(x) / den;
    const float y = weight * static_cast<float>(xd);
    write_planes(op, i, dpad, layout, y);
    const double delta = static_cast<double>(weight) * xd - static_cast<double>(__bfloat162float(__float2bfloat16_rn(y)));
    rs += delta * delta;
  }
  if (resid_out != nullptr) {
    rs = warp_sum(rs);
    if (lane == 0) resid_out[row] = __double2float_ru(rs);
  }
  for (int pl = 0; pl < planes; ++pl)
    for (int i = d + lane; i < dpad; i += 32) op[pl * dpad + i] = __float2bfloat16_rn(0.f);
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
normalize_f64_kernel(const T* __restrict__ src, int64_t n, int d, int64_t src_ld, double* __restrict__ dst,
                     int64_t dst_ld, int norm_mode) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  if (row >= n) return;
  const T* __restrict__ r = src + row * src_ld;
  double ss = 0.0;
  for (int i = lane; i < d; i += 32) {
    const double x = static_cast<double>(r[i]);
    ss += x * x;
  }
  ss = warp_sum(ss);
  double nrm = sqrt(ss);
  if (norm_mode == XMVE_NORM_EPS) nrm = fmax(nrm, 1e-12);
  double* __restrict__ o = dst + row * dst_ld;
  for (int i = lane; i < d; i += 32) o[i] = static_cast<double>(r[i]) / nrm;   // 1.0 * X / norm
}

}  // namespace
}  // namespace xmve

extern "C" int xmve_prepare_rows(const void* src, int src_dtype, int64_t n, int d, int frames, int64_t src_ld,
                                 float* raw_out, int64_t raw_ld, int64_t raw_off, double* norm_out, float* resid_out,
                                 void* op_out, int64_t op_ld, int64_t op_off, int op_layout, float weight,
                                 int norm_mode, void* stream) {
  using namespace xmve;
  XMVE_DEVICE_OR_RETURN();
  XMVE_REQUIRE(src != nullptr && n >= 0 && d > 0 && frames >= 1, "prepare_rows: bad src / n / d / frames");
  XMVE_REQUIRE(src_ld >= static_cast<int64_t>(frames) * d, "prepare_rows: src_ld < frames * d");
  XMVE_REQUIRE(src_dtype == XMVE_F32 || src_dtype == XMVE_F64, "prepare_rows: src_dtype must be f32 or f64");
  XMVE_REQUIRE(op_layout >= XMVE_OP_X1 && op_layout <= XMVE_OP_X3_CORPUS, "prepare_rows: bad op_layout");
  const int dpad = (d + 63) / 64 * 64;
  const int planes = op_layout == XMVE_OP_X1 ? 1 : 3;
  XMVE_REQUIRE(op_out == nullptr || op_ld >= op_off + static_cast<int64_t>(planes) * dpad,
               "prepare_rows: operand row too short for %d plane(s) of %d", planes, dpad);
  XMVE_REQUIRE(raw_out == nullptr || raw_ld >= raw_off + d, "prepare_rows: raw row too short");
  XMVE_REQUIRE(resid_out == nullptr || op_out != nullptr, "prepare_rows: resid_out needs op_out");
  if (n == 0) return XMVE_OK;
  const unsigned grid = static_cast<unsigned>((n + WARPS - 1) / WARPS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* op = static_cast<__nv_bfloat16*>(op_out);
  if (src_dtype == XMVE_F32) {
    const int vec4 = (frames == 1 && d % 4 == 0 && src_ld % 4 == 0 && aligned16(src) &&
                      (raw_out == nullptr || (raw_ld % 4 == 0 && raw_off % 4 == 0 && aligned16(raw_out))))
                         ? 1 : 0;
    prepare_rows_kernel<float><<<grid, WARPS * 32, 0, st>>>(static_cast<const float*>(src), n, d, frames, src_ld,
                                                            raw_out, raw_ld, raw_off, norm_out, resid_out, op, op_ld,
                                                            op_off, op_layout, dpad, weight, norm_mode, vec4);
  } else {
    prepare_rows_kernel<double><<<grid, WARPS * 32, 0, st>>>(static_cast<const double*>(src), n, d, frames, src_ld,
                                                             raw_out, raw_ld, raw_off, norm_out, resid_out, op, op_ld,
                                                             op_off, op_layout, dpad, weight, norm_mode, 0);
  }
  return launch_status("prepare_rows_kernel");
}

extern "C" int xmve_normalize_f64(const void* src, int src_dtype, int64_t n, int d, int64_t src_ld, double* dst,
                                  int64_t dst_ld, int norm_mode, void* stream) {
  using namespace xmve;
  XMVE_DEVICE_OR_RETURN();
  XMVE_REQUIRE(src != nullptr && dst != nullptr && n >= 0 && d > 0 && src_ld >= d && dst_ld >= d,
               "normalize_f64: bad arguments");
  XMVE_REQUIRE(src_dtype == XMVE_F32 || src_dtype == XMVE_F64, "normalize_f64: src_dtype must be f32 or f64");
  if (n == 0) return XMVE_OK;
  const unsigned grid = static_cast<unsigned>((n + WARPS - 1) / WARPS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_dtype == XMVE_F32)
    normalize_f64_kernel<float><<<grid, WARPS * 32, 0, st>>>(static_cast<const float*>(src), n, d, src_ld, dst,
                                                             dst_ld, norm_mode);
  else
    normalize_f64_kernel<double><<<grid, WARPS * 32, 0, st>>>(static_cast<const double*>(src), n, d, src_ld, dst,
                                                              dst_ld, norm_mode);
  return launch_status("normalize_f64_kernel");
}
