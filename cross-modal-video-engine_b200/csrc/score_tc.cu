// K2: query x corpus score kernel for sm_100a.
//
//   S[q, v] = sum_k A[q, k] * B[v, k]        A = query operand, B = corpus operand (bf16, K-major)
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B swizzle, multi-stage smem ring)
//   warp 1      MMA issuer     (tcgen05.mma, fp32 accumulators in TMEM; one elected thread)
//   warp 2      TMEM allocator (512 columns = two 256-column accumulator stages)
//   warps 4-7   epilogue       (tcgen05.ld, one query row per thread)
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// Two epilogues share the main loop:
//   STORE   out = alpha * S                                   (cal_error / sampling pass)
//   FILTER  per-row window (lo, hi]: count scores above hi, append (score, index) of scores inside
//           the window to a per-row candidate list -- the score matrix never reaches HBM.
//
// Two tile shapes:
//   PAIR (default)  two CTAs of a cluster drive one tcgen05.mma.cta_group::2 of 256 x 256 x 16: each CTA
//                   holds its 128 query rows and HALF of the 256-row corpus tile, so the shared-memory and
//                   L2 traffic per flop drop by a third against the single-CTA tile (32 KB instead of 48 KB
//                   per SM and k-block) -- the single-CTA kernel sits at the shared-memory bandwidth wall
//                   (96 B/clk of operand reads + 96 B/clk of TMA writes against 128 B/clk).
//   SINGLE          one CTA, 128 x 256 x 16 (cta_group::1); kept for comparison (XMVE_CTA_PAIR=0).
//
// Work is cut into units of ONE corpus tile x a group of query tiles, walked query-tile-major, inside
// super-blocks of the query operand sized for L2 (see plan_schedule()).
//
// Replaces np.dot(l2norm(captions), l2norm(videos).T) in LINAS-engine/evaluation.py:21,45,79 and
// P @ index.T in MultiFusion/src/validate.py:73,90, and (FILTER) the per-row argsort that follows
// them (LINAS-engine/inference.py:79, MultiFusion/src/validate.py:74,92).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace xmve {
namespace {

constexpr int BM = 128;              // query rows per CTA tile (TMEM lanes)
constexpr int BN = 256;              // corpus rows per tile (TMEM columns per accumulator stage)
constexpr int BK = 64;               // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int ACC_STAGES = 2;
constexpr int THREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int TMEM_COLS = ACC_STAGES * BN;   // 512
constexpr int A_BYTES = BM * BK * 2;         // 16 KB

enum { MODE_STORE = 0, MODE_FILTER = 1 };

template <bool PAIR>
struct Cfg {
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;            // corpus rows this CTA stages per k-block
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;        // 32 KB (pair) / 48 KB (single)
  static constexpr int STAGES = PAIR ? 6 : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int CTAS = PAIR ? 2 : 1;
  static constexpr int TILE_M = BM * CTAS;                     // query rows per MMA tile
};

struct Params {
  int64_t nq, nv;
  int k_blocks;
  int m_tiles, n_tiles, m_group, n_mgroups, sb_tiles;   // m_tiles in units of TILE_M rows
  int64_t n_units, units_per_sb;
  uint64_t hint_a, hint_b;
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
  int t, mt0, len, rot;   // corpus tile, first query tile, number of query tiles, per-worker rotation
  // i-th query tile of the unit: workers start at different query tiles so that they do not all pull the
  // same query-operand lines out of the same L2 slices at the same time
  __device__ __forceinline__ int mt(int i) const {
    const int j = i + rot;
    return mt0 + (j >= len ? j - len : j);
  }
};
// Units are ordered super-block-major: all (corpus tile, query group) units of the first `sb_tiles` query
// tiles, then the next super-block...  Within a super-block consecutive units share a corpus tile.
__device__ __forceinline__ Unit decode_unit(const Params& p, int64_t u, int worker) {
  Unit x;
  const int sb = static_cast<int>(u / p.units_per_sb);
  const int64_t r = u - static_cast<int64_t>(sb) * p.units_per_sb;
  x.t = static_cast<int>(r / p.n_mgroups);
  const int g = static_cast<int>(r - static_cast<int64_t>(x.t) * p.n_mgroups);
  const int sb_end = min(p.m_tiles, (sb + 1) * p.sb_tiles);
  x.mt0 = sb * p.sb_tiles + g * p.m_group;
  x.len = max(0, min(sb_end, x.mt0 + p.m_group) - x.mt0);
  x.rot = x.len > 0 ? worker % x.len : 0;
  return x;
}

// One 128-row x 256-column accumulator stage -> STORE or FILTER.  `row` is this thread's query row.
template <int MODE>
__device__ __forceinline__ void epilogue_tile(const Params& p, uint32_t taddr, int64_t row, int64_t col0, float lo,
                                              float hi, int& cnt) {
  const bool row_ok = row < p.nq;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr + c * 32, v);
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
}

template <int MODE, bool PAIR>
__device__ __forceinline__ void score_body(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const Params& p) {
  using C = Cfg<PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5;   // warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0;      // 0 = leader (issues the pair's MMAs)
  const int worker = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_workers = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], C::CTAS);                   // one producer arrival per CTA of the pair
      ptx::mbar_init(&empty_bar[s], 1);                        // one tcgen05.commit
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);                        // one tcgen05.commit
      ptx::mbar_init(&tempty_bar[a], 4 * C::CTAS);             // one arrival per epilogue warp (of both CTAs)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) {
      ptx::tmem_alloc_pair(tmem_slot, TMEM_COLS);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before_sync();
  if (PAIR) ptx::cluster_sync();                               // the peer's barriers exist before anyone signals them
  else __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t u = worker; u < p.n_units; u += n_workers) {
        const Unit un = decode_unit(p, u, worker);
        const int b_row = un.t * BN + static_cast<int>(rank) * C::B_ROWS;
        for (int i = 0; i < un.len; ++i) {
          const int a_row = un.mt(i) * C::TILE_M + static_cast<int>(rank) * BM;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * C::STAGE_BYTES;
            if (PAIR) {
              // both CTAs' bytes are accounted on the leader's barrier, which expects the pair's total
              if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
              else ptx::mbar_arrive_remote(&full_bar[stage], 0);
              if (p.hint_a) ptx::tma_load_2d_pair_hint(&tm_a, &full_bar[stage], sa, kb * BK, a_row, p.hint_a);
              else ptx::tma_load_2d_pair(&tm_a, &full_bar[stage], sa, kb * BK, a_row);
              if (p.hint_b) ptx::tma_load_2d_pair_hint(&tm_b, &full_bar[stage], sa + A_BYTES, kb * BK, b_row, p.hint_b);
              else ptx::tma_load_2d_pair(&tm_b, &full_bar[stage], sa + A_BYTES, kb * BK, b_row);
            } else {
              ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
              if (p.hint_a) ptx::tma_load_2d_hint(&tm_a, &full_bar[stage], sa, kb * BK, a_row, p.hint_a);
              else ptx::tma_load_2d(&tm_a, &full_bar[stage], sa, kb * BK, a_row);
              if (p.hint_b) ptx::tma_load_2d_hint(&tm_b, &full_bar[stage], sa + A_BYTES, kb * BK, b_row, p.hint_b);
              else ptx::tma_load_2d(&tm_b, &full_bar[stage], sa + A_BYTES, kb * BK, b_row);
            }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) =================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(C::TILE_M, BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int64_t u = worker; u < p.n_units; u += n_workers) {
        const Unit un = decode_unit(p, u, worker);
        for (int i = 0; i < un.len; ++i) {
          ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);     // epilogue(s) have drained this accumulator
          ptx::tc_fence_after_sync();
          const uint32_t tmem_d = tmem_base + acc * BN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);           // TMA bytes (of both CTAs) have landed
            ptx::tc_fence_after_sync();
            const uint32_t sa = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
            const uint64_t da = ptx::smem_desc_k_sw128(sa);
            const uint64_t db = ptx::smem_desc_k_sw128(sa + A_BYTES);
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
              // +32 bytes along K inside the 128-byte swizzle row = +2 in the (addr >> 4) field
              if (PAIR) ptx::umma_bf16_pair(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
              else ptx::umma_bf16(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs) when the MMAs retire
            if (PAIR) ptx::umma_commit_pair(&empty_bar[stage], 0x3);
            else ptx::umma_commit(&empty_bar[stage]);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
          if (PAIR) ptx::umma_commit_pair(&tfull_bar[acc], 0x3);  // accumulator complete -> both epilogues
          else ptx::umma_commit(&tfull_bar[acc]);
          if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ================================ epilogue ====================================
    const int quarter = warp & 3;                              // TMEM lane quarter this warp may read
    const int row_in_tile = static_cast<int>(rank) * BM + quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t u = worker; u < p.n_units; u += n_workers) {
      const Unit un = decode_unit(p, u, worker);
      const int64_t col0 = static_cast<int64_t>(un.t) * BN;
      for (int i = 0; i < un.len; ++i) {
        const int64_t row = static_cast<int64_t>(un.mt(i)) * C::TILE_M + row_in_tile;
        float lo = __int_as_float(0x7f800000), hi = __int_as_float(0x7f800000);
        int cnt = 0;
        if (MODE == MODE_FILTER && row < p.nq) {
          lo = p.lo[row];
          if (p.hi != nullptr) hi = p.hi[row];
        }
        ptx::mbar_wait(&tfull_bar[acc], acc_phase);
        ptx::tc_fence_after_sync();
        epilogue_tile<MODE>(p, tmem_base + lane_addr + acc * BN, row, col0, lo, hi, cnt);
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {                                       // one arrival per warp on the LEADER's barrier
          if (PAIR && rank != 0) ptx::mbar_arrive_remote(&tempty_bar[acc], 0);
          else ptx::mbar_arrive(&tempty_bar[acc]);
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        if (MODE == MODE_FILTER && cnt != 0 && p.count_above != nullptr) atomicAdd(&p.count_above[row], cnt);
      }
    }
  }

  ptx::tc_fence_before_sync();
  if (PAIR) ptx::cluster_sync();                               // nobody leaves while the peer may still signal it
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
score_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const Params p) {
  score_body<MODE, false>(tm_a, tm_b, p);
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
score_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const Params p) {
  score_body<MODE, true>(tm_a, tm_b, p);
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
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (const char* env = getenv("XMVE_L2PROMO")) promo = static_cast<CUtensorMapL2promotion>(atoi(env) & 3);
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
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

// L2 plan (round-1 profiles).  A line only hits if it is re-touched before L2 turns over, and the two
// dies keep their own copies of shared lines, so what must fit comfortably is
//   (query operand of the current super-block) + (corpus tiles live across the concurrent workers).
// Queries are therefore walked in super-blocks of <= ~16 MB of operand, and inside a super-block
// several workers share each corpus tile (m_group query tiles each).
void plan_schedule(Params& p, int tile_m, int k, int workers) {
  p.m_tiles = static_cast<int>((p.nq + tile_m - 1) / tile_m);
  p.n_tiles = static_cast<int>((p.nv + BN - 1) / BN);
  const int64_t a_tile_bytes = static_cast<int64_t>(tile_m) * k * 2;
  int64_t sb_mb = 16;
  if (const char* env = getenv("XMVE_SB_MB")) sb_mb = atoi(env);
  int64_t sbt = (sb_mb << 20) / a_tile_bytes;
  if (sbt < 1) sbt = 1;
  if (sbt > p.m_tiles) sbt = p.m_tiles;
  const int64_t n_sb = (p.m_tiles + sbt - 1) / sbt;
  sbt = (p.m_tiles + n_sb - 1) / n_sb;                       // balance the super-blocks
  int64_t mg = sbt / ((workers + 19) / 20);
  const int64_t balance = (static_cast<int64_t>(p.n_tiles) * sbt) / (static_cast<int64_t>(workers) * 8);
  if (mg > balance) mg = balance;                            // keep >= 8 units per worker for tail balance
  if (const char* env = getenv("XMVE_MGROUP")) mg = atoi(env);
  if (mg < 1) mg = 1;
  if (mg > sbt) mg = sbt;
  p.sb_tiles = static_cast<int>(sbt);
  p.m_group = static_cast<int>(mg);
  p.n_mgroups = static_cast<int>((sbt + mg - 1) / mg);
  p.units_per_sb = static_cast<int64_t>(p.n_tiles) * p.n_mgroups;
  p.n_units = p.units_per_sb * n_sb;
  static const uint64_t hints[3] = {0, ptx::L2_EVICT_LAST, ptx::L2_EVICT_FIRST};
  p.hint_a = p.hint_b = 0;
  if (const char* env = getenv("XMVE_HINT_A")) p.hint_a = hints[atoi(env) % 3];
  if (const char* env = getenv("XMVE_HINT_B")) p.hint_b = hints[atoi(env) % 3];
}

template <int MODE, bool PAIR>
int launch_cfg(const void* a_op, int64_t nq, int64_t a_ld, const void* b_op, int64_t nv, int64_t b_ld,
               int64_t b_row_step, int k, Params p, cudaStream_t stream) {
  using C = Cfg<PAIR>;
  CUtensorMap tm_a, tm_b;
  int s = make_operand_map(&tm_a, a_op, nq, k, a_ld, BM);
  if (s != XMVE_OK) return s;
  s = make_operand_map(&tm_b, b_op, nv, k, b_ld * b_row_step, C::B_ROWS);
  if (s != XMVE_OK) return s;
  const int sms = sm_count();
  if (sms <= 0) return fail(XMVE_ERR_DEVICE, "score: cannot query the SM count");
  const int workers_max = sms / C::CTAS;
  p.nq = nq;
  p.nv = nv;
  p.k_blocks = k / BK;
  plan_schedule(p, C::TILE_M, k, workers_max);
  const int workers = static_cast<int>(p.n_units < workers_max ? p.n_units : workers_max);
  const int grid = workers * C::CTAS;
  static bool attr_set = false;                               // one flag per <MODE, PAIR> instantiation
  if (PAIR) {
    if (!attr_set) {
      XMVE_CUDA(cudaFuncSetAttribute(score_pair_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
      attr_set = true;
    }
    score_pair_kernel<MODE><<<grid, THREADS, C::SMEM_BYTES, stream>>>(tm_a, tm_b, p);
  } else {
    if (!attr_set) {
      XMVE_CUDA(cudaFuncSetAttribute(score_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
      attr_set = true;
    }
    score_kernel<MODE><<<grid, THREADS, C::SMEM_BYTES, stream>>>(tm_a, tm_b, p);
  }
  return launch_status(PAIR ? "score_pair_kernel" : "score_kernel");
}

template <int MODE>
int launch(const void* a_op, int64_t nq, int64_t a_ld, const void* b_op, int64_t nv, int64_t b_ld, int64_t b_row_step,
           int k, Params p, cudaStream_t stream) {
  bool pair = true;
  if (const char* env = getenv("XMVE_CTA_PAIR")) pair = atoi(env) != 0;
  if (pair) return launch_cfg<MODE, true>(a_op, nq, a_ld, b_op, nv, b_ld, b_row_step, k, p, stream);
  return launch_cfg<MODE, false>(a_op, nq, a_ld, b_op, nv, b_ld, b_row_step, k, p, stream);
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
