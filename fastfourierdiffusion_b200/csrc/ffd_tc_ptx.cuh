// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the tensor-core kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ffd_internal.h"

// 16-bit operand format of every tensor-core GEMM of this library (weights, token tiles, Q / K tiles, relu(H), the
// out-projection operand; fp32 accumulation throughout): 1 = fp16, 0 = bf16.  Same tensor throughput, but fp16 carries
// an 11-bit mantissa against bf16's 8: on the well-conditioned ECG trace (tests: test_ecg_ve_well_conditioned_trace)
// bf16 operands leave a score error of 9.5e-3 (max-norm relative), above the 5e-3 the path is held to.  Range is the
// price: every fp32 -> fp16 conversion SATURATES at +-65504 instead of producing infinities (activations of the
// sampling path are LayerNorm outputs and projections of them, |.| < 1e3; the un-normalised embedding of a
// diverging random-init run reaches ~5e2).  P and V were fp16 already.
#ifndef FFD_TC_FP16
#define FFD_TC_FP16 1
#endif

namespace ffd {
namespace tc {

#if FFD_TC_FP16
using op16 = __half;
__host__ __device__ inline op16 to_op(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }
__host__ __device__ inline float op_to_float(op16 v) { return __half2float(v); }
__device__ __forceinline__ uint16_t op_bits(float v) { return __half_as_ushort(to_op(v)); }
#else
using op16 = __nv_bfloat16;
__host__ __device__ inline op16 to_op(float v) { return __float2bfloat16_rn(v); }
__host__ __device__ inline float op_to_float(op16 v) { return __bfloat162float(v); }
__device__ __forceinline__ uint16_t op_bits(float v) { return __bfloat16_as_ushort(to_op(v)); }
#endif
// v rounded to the operand format (the "hi" part of a bias carried as hi + lo against two constant-1 columns)
__host__ __device__ inline float round_op(float v) { return op_to_float(to_op(v)); }

constexpr int TILE_M = 128;   // rows per token tile (UMMA M)

__host__ __device__ constexpr int dp_of(int d) { return (d + 2 + 15) / 16 * 16; }

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)  // suspend-time hint (ns): sleep in hardware, do not spin
      : "memory");
  return ok != 0;
}
// Bounded wait (-DFFD_WAIT_TRAPS=1, the development build): a protocol bug traps (launch error) instead of hanging
// the GPU.  try_wait suspends the thread in hardware for a while before returning false, so 2^24 failed polls is
// seconds of waiting.  The release build polls without the counter: the counter line was 29 % of the warp
// instructions the FFN kernel executed (round-1 ncu), issue slots the softmax warps of the attention kernel need.
#ifndef FFD_WAIT_TRAPS
#define FFD_WAIT_TRAPS 0
#endif
#if FFD_WAIT_TRAPS
#define FFD_POLL_GUARD(polls) do { if (++(polls) > (1u << 24)) __trap(); } while (0)
#else
#define FFD_POLL_GUARD(polls) do { (void)(polls); } while (0)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) FFD_POLL_GUARD(polls);
}
// Same for the roles that wait long and are not latency-critical (producers, tile warps): sleep between polls so
// that they do not compete with the softmax warps for issue slots (the hardware wakes a suspended try_wait on
// every barrier event of the CTA, so plain try_wait loops poll at a high rate).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t sleep_ns) {
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(sleep_ns);
    FFD_POLL_GUARD(polls);
  }
}
// Wait for several mbarriers AT ONCE: every lane polls its own barrier (shared-memory address + parity; lanes without
// a barrier of their own pass a copy of another lane's).  A try_wait round trip costs ~150 cycles even when the phase
// is already complete, and an MMA-issuing warp pays them back to back -- measured: three waits in a row were half
// of an issue loop that has to turn around in less than the tensor pipe needs for the MMAs it issues.  Safe as long
// as none of the barriers can complete a SECOND phase before this warp moves on (each lane latches its result).
__device__ __forceinline__ void mbar_wait_lanes(uint32_t bar_addr, uint32_t parity) {
  uint32_t done = 0, polls = 0;
  for (;;) {
    if (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar_addr), "r"(parity), "r"(20000u)
          : "memory");
    }
    if (__all_sync(0xffffffffu, done != 0)) break;
    FFD_POLL_GUARD(polls);
  }
}

__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 1-D bulk async copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// bulk async store shared -> global (SASS: UBLKCP), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the shared-memory sources of all committed bulk stores have been read
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// bar.sync after which v is re-read from shared memory (volatile): arithmetic that depends on the returned value
// cannot be scheduled above the barrier by ptxas (a memory clobber orders memory operations only, and an operand the
// asm text does not mention is no dependency at the PTX level).
__device__ __forceinline__ float named_bar_sync_reload(int id, int nthreads, const float* v) {
  float r;
  asm volatile("bar.sync %1, %2;\n\tld.volatile.shared.f32 %0, [%3];"
               : "=f"(r) : "r"(id), "r"(nthreads), "r"(smem_u32(v)) : "memory");
  return r;
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor, K-major, no swizzle: 8x(16 B) core matrices; LBO = byte step
// between the two K halves of one MMA (k-groups), SBO = byte step between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32, A/B in the operand format (fp16: code 0, bf16: code 1), both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((FFD_TC_FP16 ? 0u : 1u) << 7) | ((FFD_TC_FP16 ? 0u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// same with fp16 A/B operands (format code 0)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// The MMA-issuing warp runs CONVERGED: every lane executes the issue loop with warp-uniform operands and
// the single issuing lane is elected inside the asm statement.  (Issuing from inside an `if (lane == 0)`
// branch makes ptxas wrap every UTCHMMA in a per-thread "waterfall" loop with R2UR moves -- measured ~240
// cycles per MMA instead of ~20.)
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on an mbarrier once every MMA issued so far by the elected lane has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

#define FFD_TMEM_LD16(taddr, r)                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(taddr))
#define FFD_TMEM_LD8(taddr, r)                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"             \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
               : "r"(taddr))
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// relu + round-to-nearest (saturating for fp16) conversion of two floats to the operand format in one instruction;
// lo -> low half (lower k index)
__device__ __forceinline__ uint32_t relu_pack_op(uint32_t lo_bits, uint32_t hi_bits) {
  uint32_t r;
#if FFD_TC_FP16
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
#else
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
#endif
  return r;
}
__device__ __forceinline__ uint32_t pack_op(float lo, float hi) {
  uint32_t r;
#if FFD_TC_FP16
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
#else
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
#endif
  return r;
}


// A operand read from TMEM (16-bit pairs packed per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// ---- chained issue: several K-steps of one GEMM under a single election ------------------------------
// Measured on B200 (tools/mma_bench.cu): one tcgen05.mma per asm statement costs ~92 (A in smem) / ~110 (A in
// TMEM) cycles whatever N is; eight per statement cost max(N/2, ~25) cycles each with A in TMEM and
// max(N/2, ~42) with A in shared memory -- so K-loops are issued as chains.
#define FFD_TS_HEAD                                                                   \
  "{\n\t.reg .pred q, p;\n\t.reg .b32 ra;\n\t.reg .b64 bd, bs;\n\t"                     \
  "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"                          \
  "mov.b32 ra, %1;\n\tmov.b64 bd, %2;\n\tcvt.u64.u32 bs, %6;\n\t"                      \
  "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ra], bd, %3, p;\n\t"
#define FFD_TS_NEXT \
  "add.u32 ra, ra, %5;\n\tadd.u64 bd, bd, bs;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ra], bd, %3, 1;\n\t"
#define FFD_TS_ARGS ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc), "r"(a_step), "r"(b_step) : "memory"

// D[tmem d] (+)= A[tmem a + i*a_step columns] . B[desc bd + i*b_step (16-byte units)], i < NK
template <int NK>
__device__ __forceinline__ void umma_ts_chain(uint32_t d, uint32_t a, uint32_t a_step, uint64_t bd, uint32_t b_step,
                                              uint32_t idesc, uint32_t acc) {
  if constexpr (NK >= 8) {
    asm volatile(FFD_TS_HEAD FFD_TS_NEXT FFD_TS_NEXT FFD_TS_NEXT FFD_TS_NEXT FFD_TS_NEXT FFD_TS_NEXT FFD_TS_NEXT
                 "}" FFD_TS_ARGS);
    umma_ts_chain<NK - 8>(d, a + 8 * a_step, a_step, bd + 8 * (uint64_t)b_step, b_step, idesc, 1);
  } else if constexpr (NK >= 4) {
    asm volatile(FFD_TS_HEAD FFD_TS_NEXT FFD_TS_NEXT FFD_TS_NEXT "}" FFD_TS_ARGS);
    umma_ts_chain<NK - 4>(d, a + 4 * a_step, a_step, bd + 4 * (uint64_t)b_step, b_step, idesc, 1);
  } else if constexpr (NK >= 2) {
    asm volatile(FFD_TS_HEAD FFD_TS_NEXT "}" FFD_TS_ARGS);
    umma_ts_chain<NK - 2>(d, a + 2 * a_step, a_step, bd + 2 * (uint64_t)b_step, b_step, idesc, 1);
  } else if constexpr (NK == 1) {
    asm volatile(FFD_TS_HEAD "}" FFD_TS_ARGS);
  }
}

// Four independent 3-step chains under ONE election (the split-K P.V of the attention kernel):
//   D_q[tmem d + 16 q] = sum_j A[tmem a + q*a_qstride + 8 j] . B[desc bd + (3 q + j) * b_step],  q < 4, j < 3
__device__ __forceinline__ void umma_ts_4x3(uint32_t d, uint32_t a, uint32_t a_qstride, uint64_t bd, uint32_t b_step,
                                            uint32_t idesc) {
#define FFD_Q3(first)                                                                               \
  "@q tcgen05.mma.cta_group::1.kind::f16 [rd], [ra], bd, %3, 0;\n\t"                                  \
  "add.u32 rt, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t"                                                     \
  "@q tcgen05.mma.cta_group::1.kind::f16 [rd], [rt], bd, %3, 1;\n\t"                                  \
  "add.u32 rt, ra, 16;\n\tadd.u64 bd, bd, bs;\n\t"                                                    \
  "@q tcgen05.mma.cta_group::1.kind::f16 [rd], [rt], bd, %3, 1;\n\t"                                  \
  "add.u32 rd, rd, 16;\n\tadd.u32 ra, ra, %4;\n\tadd.u64 bd, bd, bs;\n\t"
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b32 rd, ra, rt;\n\t.reg .b64 bd, bs;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b32 rd, %0;\n\tmov.b32 ra, %1;\n\tmov.b64 bd, %2;\n\tcvt.u64.u32 bs, %5;\n\t" FFD_Q3(1) FFD_Q3(0) FFD_Q3(0)
          FFD_Q3(0) "}" ::"r"(d),
      "r"(a), "l"(bd), "r"(idesc), "r"(a_qstride), "r"(b_step)
      : "memory");
#undef FFD_Q3
}


// ---- CTA pairs (cta_group::2): two SMs of a cluster run one M = 256 MMA; each supplies half of the B operand --------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA 0 of the cluster (the MMA-issuing CTA).  Default
// semantics (as CUTLASS's ClusterBarrier::arrive): what is handed over lives in TMEM / is ordered by the tcgen05
// fences, and a cluster-scope release would cost a memory fence per arriving thread (measured: 4x slower epilogue).
__device__ __forceinline__ void mbar_arrive_cta0(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
// wait on a barrier that threads of the PEER CTA arrive on (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    FFD_POLL_GUARD(polls);
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b16 m, 3;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(
          smem_u32(bar))
      : "memory");
}
#define FFD_TS2_HEAD                                                                  \
  "{\n\t.reg .pred q, p;\n\t.reg .b32 ra;\n\t.reg .b64 bd, bs;\n\t"                     \
  "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"                          \
  "mov.b32 ra, %1;\n\tmov.b64 bd, %2;\n\tcvt.u64.u32 bs, %6;\n\t"                      \
  "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [ra], bd, %3, p;\n\t"
#define FFD_TS2_NEXT \
  "add.u32 ra, ra, %5;\n\tadd.u64 bd, bd, bs;\n\t@q tcgen05.mma.cta_group::2.kind::f16 [%0], [ra], bd, %3, 1;\n\t"
// D[tmem d] (+)= A[tmem a + i*a_step columns] . B[desc bd + i*b_step], i < NK, over the CTA pair (M = 256)
template <int NK>
__device__ __forceinline__ void umma2_ts_chain(uint32_t d, uint32_t a, uint32_t a_step, uint64_t bd, uint32_t b_step,
                                               uint32_t idesc, uint32_t acc) {
  if constexpr (NK >= 4) {
    asm volatile(FFD_TS2_HEAD FFD_TS2_NEXT FFD_TS2_NEXT FFD_TS2_NEXT "}" FFD_TS_ARGS);
    umma2_ts_chain<NK - 4>(d, a + 4 * a_step, a_step, bd + 4 * (uint64_t)b_step, b_step, idesc, 1);
  } else if constexpr (NK >= 2) {
    asm volatile(FFD_TS2_HEAD FFD_TS2_NEXT "}" FFD_TS_ARGS);
    umma2_ts_chain<NK - 2>(d, a + 2 * a_step, a_step, bd + 2 * (uint64_t)b_step, b_step, idesc, 1);
  } else if constexpr (NK == 1) {
    asm volatile(FFD_TS2_HEAD "}" FFD_TS_ARGS);
  }
}

#define FFD_SS_HEAD                                                                   \
  "{\n\t.reg .pred q, p;\n\t.reg .b64 ad, bd, as, bs;\n\t"                             \
  "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"                          \
  "mov.b64 ad, %1;\n\tmov.b64 bd, %2;\n\tcvt.u64.u32 as, %5;\n\tcvt.u64.u32 bs, %6;\n\t" \
  "@q tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %3, p;\n\t"
#define FFD_SS_NEXT \
  "add.u64 ad, ad, as;\n\tadd.u64 bd, bd, bs;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %3, 1;\n\t"
#define FFD_SS_ARGS ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(a_step), "r"(b_step) : "memory"

// same with both operands in shared memory; steps in 16-byte descriptor units
template <int NK>
__device__ __forceinline__ void umma_ss_chain(uint32_t d, uint64_t ad, uint32_t a_step, uint64_t bd, uint32_t b_step,
                                              uint32_t idesc, uint32_t acc) {
  if constexpr (NK >= 4) {
    asm volatile(FFD_SS_HEAD FFD_SS_NEXT FFD_SS_NEXT FFD_SS_NEXT "}" FFD_SS_ARGS);
    umma_ss_chain<NK - 4>(d, ad + 4 * (uint64_t)a_step, a_step, bd + 4 * (uint64_t)b_step, b_step, idesc, 1);
  } else if constexpr (NK >= 2) {
    asm volatile(FFD_SS_HEAD FFD_SS_NEXT "}" FFD_SS_ARGS);
    umma_ss_chain<NK - 2>(d, ad + 2 * (uint64_t)a_step, a_step, bd + 2 * (uint64_t)b_step, b_step, idesc, 1);
  } else if constexpr (NK == 1) {
    asm volatile(FFD_SS_HEAD "}" FFD_SS_ARGS);
  }
}

// ---- whole issue steps of the fused layer kernel under ONE election (ffd_tc_layer.cu) -------------------------
// FFN step:  Y += relu(H_b)[tmem a2 + 8 i] . W2[desc w2 + i * s2], i < 4   (accumulating)
//            commit -> e2 (W2 slot free)  [, commit -> yf if `last`]
//            if `more`:  H_b = X[tmem x + 8 i] . W1[desc w1 + i * s1], i < 5;  commit -> hf (H_b ready), e1 (W1 slot free)
#ifdef FFD_L_NO_FFN_MMA
#define FFD_FQ "qx"
#define FFD_FQM "qx"
#else
#define FFD_FQ "q"
#define FFD_FQM "qm"
#endif
#ifdef FFD_L_NO_ATTN_MMA
#define FFD_AE "ex"
#define FFD_AEM "ex"
#else
#define FFD_AE "e"
#define FFD_AEM "em"
#endif
__device__ __forceinline__ void ffn_issue_step(uint32_t y, uint32_t a2, uint64_t w2, uint32_t s2, uint32_t idesc2,
                                               uint32_t e2, uint32_t yf, uint32_t last, uint32_t more, uint32_t h,
                                               uint32_t x, uint64_t w1, uint32_t s1, uint32_t idesc1, uint32_t hf,
                                               uint32_t e1) {
  asm volatile(
      "{\n\t.reg .pred q, ql, qm, qx;\n\tsetp.ne.b32 qx, 0, 0;\n\t.reg .b32 ra;\n\t.reg .b64 bd, bs;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.and.b32 ql, %7, 0, q;\n\t"
      "setp.ne.and.b32 qm, %8, 0, q;\n\t"
      "mov.b32 ra, %1;\n\tmov.b64 bd, %2;\n\tcvt.u64.u32 bs, %3;\n\t"
      "@" FFD_FQ " tcgen05.mma.cta_group::1.kind::f16 [%0], [ra], bd, %4, 1;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_FQ " tcgen05.mma.cta_group::1.kind::f16 [%0], [ra], bd, %4, 1;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_FQ " tcgen05.mma.cta_group::1.kind::f16 [%0], [ra], bd, %4, 1;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_FQ " tcgen05.mma.cta_group::1.kind::f16 [%0], [ra], bd, %4, 1;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
      "@ql tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "mov.b32 ra, %10;\n\tmov.b64 bd, %11;\n\tcvt.u64.u32 bs, %12;\n\t"
      "@" FFD_FQM " tcgen05.mma.cta_group::1.kind::f16 [%9], [ra], bd, %13, 0;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_FQM " tcgen05.mma.cta_group::1.kind::f16 [%9], [ra], bd, %13, 1;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_FQM " tcgen05.mma.cta_group::1.kind::f16 [%9], [ra], bd, %13, 1;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_FQM " tcgen05.mma.cta_group::1.kind::f16 [%9], [ra], bd, %13, 1;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_FQM " tcgen05.mma.cta_group::1.kind::f16 [%9], [ra], bd, %13, 1;\n\t"
      "@qm tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%14];\n\t"
      "@qm tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%15];\n\t}" ::"r"(y),
      "r"(a2), "l"(w2), "r"(s2), "r"(idesc2), "r"(e2), "r"(yf), "r"(last), "r"(more), "r"(h), "r"(x), "l"(w1), "r"(s1),
      "r"(idesc1), "r"(hf), "r"(e1)
      : "memory");
}
// Attention step:  O_g = P[tmem p + 48 g + 8 j] . V[desc v + (3 g + j) * sv], g < 2, j < 3   (two accumulators, o and o + 16)
//                  commit -> of (O ready), ve (V slot free)
//                  if `more`:  S = Q[desc q] . K[desc k] into tmem sd;  commit -> sf (scores ready), ke (K slot free)
__device__ __forceinline__ void attn_issue_step(uint32_t o, uint32_t p, uint64_t v, uint32_t sv, uint32_t idesc_pv,
                                                uint32_t of, uint32_t ve, uint32_t more, uint32_t sd, uint64_t q,
                                                uint64_t k, uint32_t idesc_s, uint32_t sf, uint32_t ke) {
  asm volatile(
      "{\n\t.reg .pred e, em, ex;\n\tsetp.ne.b32 ex, 0, 0;\n\t.reg .b32 ra, rd;\n\t.reg .b64 bd, bs;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.and.b32 em, %7, 0, e;\n\t"
      "mov.b32 ra, %1;\n\tmov.b32 rd, %0;\n\tmov.b64 bd, %2;\n\tcvt.u64.u32 bs, %3;\n\t"
      "@" FFD_AE " tcgen05.mma.cta_group::1.kind::f16 [rd], [ra], bd, %4, 0;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_AE " tcgen05.mma.cta_group::1.kind::f16 [rd], [ra], bd, %4, 1;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_AE " tcgen05.mma.cta_group::1.kind::f16 [rd], [ra], bd, %4, 1;\n\t"
      "add.u32 ra, ra, 32;\n\tadd.u32 rd, rd, 16;\n\tadd.u64 bd, bd, bs;\n\t"
      "@" FFD_AE " tcgen05.mma.cta_group::1.kind::f16 [rd], [ra], bd, %4, 0;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_AE " tcgen05.mma.cta_group::1.kind::f16 [rd], [ra], bd, %4, 1;\n\t"
      "add.u32 ra, ra, 8;\n\tadd.u64 bd, bd, bs;\n\t@" FFD_AE " tcgen05.mma.cta_group::1.kind::f16 [rd], [ra], bd, %4, 1;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "@" FFD_AEM " tcgen05.mma.cta_group::1.kind::f16 [%8], %9, %10, %11, 0;\n\t"
      "@em tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%12];\n\t"
      "@em tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%13];\n\t}" ::"r"(o),
      "r"(p), "l"(v), "r"(sv), "r"(idesc_pv), "r"(of), "r"(ve), "r"(more), "r"(sd), "l"(q), "l"(k), "r"(idesc_s),
      "r"(sf), "r"(ke)
      : "memory");
}

#define FFD_TMEM_ST16(taddr, r)                                                                                     \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),  \
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])         \
               : "memory")
#define FFD_TMEM_ST8(taddr, r)                                                                        \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),  \
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])   \
               : "memory")
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2(x) for x <= 0 and 1/x for x > 0 on the FMA/ALU pipes only.  The softmax warps keep the MUFU queue of
// their sub-partition full (hundreds of cycles deep), so the few transcendental ops of the O read-out must not
// queue behind it.  exp2: round-to-nearest split x = n + f, degree-5 polynomial for 2^f on [-0.5, 0.5] (rel.
// error < 2e-7), exponent added through the integer view.  rcp: bit-trick seed + 3 Newton steps (full fp32).
__device__ __forceinline__ float exp2_fma(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;             // 1.5 * 2^23: low mantissa bits now hold round(x)
  const float f = x - (t - 12582912.f);
  float p = 1.3333558e-3f;
  p = fmaf(p, f, 9.6181291e-3f);
  p = fmaf(p, f, 5.5504109e-2f);
  p = fmaf(p, f, 2.4022651e-1f);
  p = fmaf(p, f, 6.9314718e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ float rcp_fma(float x) {
  float r = __int_as_float(0x7EF311C7 - __float_as_int(x));
  r = r * fmaf(-x, r, 2.f);
  r = r * fmaf(-x, r, 2.f);
  r = r * fmaf(-x, r, 2.f);
  r = r * fmaf(-x, r, 2.f);
  return r;
}

// exp2(x), x <= 0, on the FMA pipe with a degree-3 polynomial (rel. error 1.4e-4: below fp16's half ulp of 4.9e-4).
// The attention softmax is bound by MUFU.EX2 (16 lanes/clk/SM); evaluating a share of its exponentials here moves
// work to the otherwise idle FMA lanes.
__device__ __forceinline__ float exp2_poly3(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);          // f in [-0.5, 0.5]
  float p = fmaf(5.5029270e-2f, f, 2.4225698e-1f);
  p = fmaf(p, f, 6.9325305e-1f);
  p = fmaf(p, f, 9.9995134e-1f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t ex2_pack_f16_poly(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(exp2_poly3(hi)), "f"(exp2_poly3(lo)));
  return r;
}
// two exponentials packed as an fp16 pair; lo -> low half (lower k index)
__device__ __forceinline__ uint32_t ex2_pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(ex2_approx(hi)), "f"(ex2_approx(lo)));
  return r;
}
// exp2(lo + nm), exp2(hi + nm) packed as an fp16 pair; the two additions are ONE packed fp32x2 instruction
// (FADD2 on sm_100), which matters because the softmax loop is bound by issue slots as much as by MUFU
__device__ __forceinline__ uint32_t ex2_pack_f16_shift(float lo, float hi, float nm) {
  uint64_t x, mm, y;
  float a, b;
  uint32_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(lo), "f"(hi));
  asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(nm));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(y) : "l"(x), "l"(mm));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(y));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(ex2_approx(b)), "f"(ex2_approx(a)));
  return r;
}
// The same pair on the FMA pipe: round-to-nearest split x = n + f, degree-3 polynomial for 2^f on [-0.5, 0.5]
// (rel. error 1.4e-4, below fp16's half ulp of 4.9e-4), exponent added through the integer view; packed fp32x2
// arithmetic throughout.  Moves a share of the softmax exponentials off the MUFU pipe (16 lanes/clk/SM).
__device__ __forceinline__ uint32_t ex2_pack_f16_shift_poly(float lo, float hi, float nm) {
  uint64_t x, y, t, u, f, pp;
  float a, b;
  uint32_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(lo), "f"(hi));
  asm("{\n\t.reg .b64 m;\n\tmov.b64 m, {%2, %2};\n\tadd.rn.f32x2 %0, %1, m;\n\t}" : "=l"(y) : "l"(x), "f"(nm));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(y));
  a = fmaxf(a, -125.f);
  b = fmaxf(b, -125.f);
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(a), "f"(b));
  asm("{\n\t.reg .b64 m;\n\tmov.b64 m, {%2, %2};\n\tadd.rn.f32x2 %0, %1, m;\n\t}" : "=l"(t) : "l"(y), "f"(12582912.f));
  asm("{\n\t.reg .b64 m;\n\tmov.b64 m, {%2, %2};\n\tadd.rn.f32x2 %0, %1, m;\n\t}" : "=l"(u) : "l"(t), "f"(-12582912.f));
  asm("{\n\t.reg .b64 m;\n\tmov.b64 m, {%3, %3};\n\tfma.rn.f32x2 %0, %1, m, %2;\n\t}" : "=l"(f) : "l"(u), "l"(y), "f"(-1.f));
  asm("{\n\t.reg .b64 c3, c2;\n\tmov.b64 c3, {%2, %2};\n\tmov.b64 c2, {%3, %3};\n\tfma.rn.f32x2 %0, c3, %1, c2;\n\t}"
      : "=l"(pp) : "l"(f), "f"(5.5029270e-2f), "f"(2.4225698e-1f));
  asm("{\n\t.reg .b64 c;\n\tmov.b64 c, {%2, %2};\n\tfma.rn.f32x2 %0, %0, %1, c;\n\t}" : "+l"(pp) : "l"(f), "f"(6.9325305e-1f));
  asm("{\n\t.reg .b64 c;\n\tmov.b64 c, {%2, %2};\n\tfma.rn.f32x2 %0, %0, %1, c;\n\t}" : "+l"(pp) : "l"(f), "f"(9.9995134e-1f));
  float p0, p1, t0, t1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(p0), "=f"(p1) : "l"(pp));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
  const float e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  const float e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
  return r;
}
// d = a * w + c on two fp32 lanes at once (FFMA2)
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float w, float c0, float c1) {
  uint64_t a, ww, c, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(ww), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}

}  // namespace tc
}  // namespace ffd
