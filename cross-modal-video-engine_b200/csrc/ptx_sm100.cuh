// Inline-PTX wrappers for the sm_100a features the score kernel uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences).
#pragma once
#include <stdint.h>

namespace xmve {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel, never as a hung GPU.  The bound is WALL TIME
// (%globaltimer, looked at every 2^16 polls), not a spin count: under compute-sanitizer, an ncu replay or a debugger
// a legitimate wait can take orders of magnitude more polls.  -DXMVE_MBAR_TIMEOUT_S=0 disables the bound.
#ifndef XMVE_MBAR_TIMEOUT_S
#define XMVE_MBAR_TIMEOUT_S 20
#endif
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __noinline__ void mbar_timeout_check(uint64_t& t0, const char* what) {
  if (XMVE_MBAR_TIMEOUT_S == 0) return;
  const uint64_t now = global_timer_ns();
  if (t0 == 0) t0 = now;
  else if (now - t0 > static_cast<uint64_t>(XMVE_MBAR_TIMEOUT_S) * 1000000000ull) {
    printf("xmve: %s wait exceeded %d s (block %d thread %d)\n", what, XMVE_MBAR_TIMEOUT_S, blockIdx.x, threadIdx.x);
    __trap();
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFFu) == 0) mbar_timeout_check(t0, "mbarrier");
  }
}

// ---- TMA ---------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
// 2-D tiled load global -> shared, completion signalled on `bar` (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(const void* desc, uint64_t* bar, void* smem_dst, int32_t c0,
                                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Same, with an L2 eviction-priority hint (a createpolicy-encoded 64-bit word).
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_hint(const void* desc, uint64_t* bar, void* smem_dst, int32_t c0,
                                                 int32_t c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "l"(hint)
      : "memory");
}

// ---- tcgen05 -----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on `bar` once every MMA issued so far by this thread has completed (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants -------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the barrier at the same smem offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      :
      : "r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// Cluster-scope variants for handing a value to the peer CTA through its shared memory: store + release-arrive on
// the peer's barrier, acquire-wait on the local one.
__device__ __forceinline__ void st_remote_u64(void* local_addr, uint32_t cta, uint64_t v) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.u64 [ra], %2;\n\t}"
      :
      : "r"(smem_u32(local_addr)), "r"(cta), "l"(v)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      :
      : "r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  uint64_t t0 = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && (++spins & 0xFFFFu) == 0) mbar_timeout_check(t0, "cluster mbarrier");
  }
}
// TMA load issued by either CTA of a pair; the bytes are accounted on the LEADER CTA's barrier
// (bit 24 of the shared::cluster address selects the peer: clear it).
__device__ __forceinline__ void tma_load_2d_pair(const void* desc, uint64_t* bar, void* smem_dst, int32_t c0,
                                                 int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0),
        "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(const void* desc, uint64_t* bar, void* smem_dst, int32_t c0,
                                                      int32_t c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0),
        "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 256 x N x 16 MMA over the CTA pair (each CTA supplies its 128 rows of A and its half of B); issued by
// one thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same, executed by ALL lanes of a warp in uniform control flow: one elected lane issues.  With uniform operands
// the descriptors stay in uniform registers; issued from inside an `if (lane == 0)` region every operand went through
// an ELECT / R2UR.BROADCAST waterfall (~25 SASS instructions per MMA, 536 clocks per k-block against the 512 the
// tensor pipe needs: the issuing thread, not the pipe, set the pace -- profiles/r2_summary.md section 9).
__device__ __forceinline__ void umma_bf16_pair_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One 64-wide k-block = four K = 16 MMAs on consecutive 32-byte slices of the swizzled operand rows (+2 in the
// descriptors' start-address field), issued in ONE statement: one election, and the descriptors are handed over as
// their low words only -- the high word (SBO = 1024 B, version 1, SWIZZLE_128B) is the constant DESC_HI -- so that
// each MMA costs two register-to-uniform moves instead of five.  `first` = 0 overwrites the accumulator in the first
// MMA (start of a tile).  CTA_GROUP is 1 or 2.
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
template <int CTA_GROUP>
__device__ __forceinline__ void umma_bf16_kblock_warp(uint32_t tmem_d, uint32_t desc_a_lo, uint32_t desc_b_lo,
                                                      uint32_t idesc, uint32_t first) {
  static_assert(CTA_GROUP == 1 || CTA_GROUP == 2, "cta_group is 1 or 2");
  if (CTA_GROUP == 2) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\t.reg .b32 a1, a2, a3, b1, b2, b3;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "add.u32 a1, %1, 2;\n\tadd.u32 a2, %1, 4;\n\tadd.u32 a3, %1, 6;\n\t"
        "add.u32 b1, %2, 2;\n\tadd.u32 b2, %2, 4;\n\tadd.u32 b3, %2, 6;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t"
        "mov.b64 da, {a1, %5};\n\tmov.b64 db, {b1, %5};\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n\t"
        "mov.b64 da, {a2, %5};\n\tmov.b64 db, {b2, %5};\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n\t"
        "mov.b64 da, {a3, %5};\n\tmov.b64 db, {b3, %5};\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n\t}"
        :
        : "r"(tmem_d), "r"(desc_a_lo), "r"(desc_b_lo), "r"(idesc), "r"(first), "n"(DESC_HI)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\t.reg .b32 a1, a2, a3, b1, b2, b3;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "add.u32 a1, %1, 2;\n\tadd.u32 a2, %1, 4;\n\tadd.u32 a3, %1, 6;\n\t"
        "add.u32 b1, %2, 2;\n\tadd.u32 b2, %2, 4;\n\tadd.u32 b3, %2, 6;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "mov.b64 da, {a1, %5};\n\tmov.b64 db, {b1, %5};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t"
        "mov.b64 da, {a2, %5};\n\tmov.b64 db, {b2, %5};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t"
        "mov.b64 da, {a3, %5};\n\tmov.b64 db, {b3, %5};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t}"
        :
        : "r"(tmem_d), "r"(desc_a_lo), "r"(desc_b_lo), "r"(idesc), "r"(first), "n"(DESC_HI)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit_pair_warp(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
// Arrive on the barrier at this smem offset in every CTA of `mask` once the pair's MMAs have retired.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// ---- descriptors -------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 bytes with the
// 128-byte swizzle (what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups are 1024 B apart.
//   [0,14) start >> 4 | [16,30) LBO >> 4 (ignored for swizzled K-major) | [32,46) SBO >> 4 = 64
//   [46,48) version = 1 (sm_100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor, kind::f16: fp32 accumulate, bf16 A and B, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t idesc_bf16_f32(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

}  // namespace ptx
}  // namespace xmve
