// tcgen05 (5th-generation tensor core) implicit GEMM for the conv/linear contractions of the U-Net,
// fp32 in / fp32 out with the 3xTF32 split so results stay at fp32 accuracy (DESIGN.md §4 "Precision"):
//
//     a = a_hi + a_lo (both exactly representable in TF32, round-to-nearest),   w = w_hi + w_lo
//     acc += a_lo*w_hi + a_hi*w_lo + a_hi*w_hi          (fp32 accumulation in TMEM, dropped term ~2^-22)
//
// Structure (one 128 x BN output tile per CTA, optional split-K over blockIdx.z):
//   warp 0      TMA producer : per k-step (32 fp32 = one 128-byte swizzle row) loads the raw A tile through a
//                              3-D tensor map (k, l, b) -- the conv k=3 halo is the TMA out-of-bounds zero fill
//                              on the l axis, so no im2col / padding copy exists -- plus the pre-split W_hi / W_lo
//                              tiles; completion on an mbarrier (complete_tx).
//   warps 4-7   converter    : split the raw A tile into a_hi (in place) and a_lo (cvt.rna.tf32), then
//                              fence.proxy.async and signal the MMA warp.  Elementwise on smem addresses, so it is
//                              independent of the 128B swizzle pattern.
//   warp 1      MMA issuer   : one elected thread issues 12 tcgen05.mma.kind::tf32 (M128 x BN x K8) per k-step
//                              from shared-memory descriptors (K-major, SWIZZLE_128B); tcgen05.commit releases the
//                              stage to the producer and, after the last k-step, hands the accumulator to the epilogue.
//   warp 2      TMEM allocator (BN fp32 columns x 128 lanes).
//   warps 4-7   epilogue     : tcgen05.ld 32x32b (thread = one output row, 32 columns at a time) -> bias /
//                              time-embedding row / SiLU / GELU / GEGLU / GLU / residual -> 128-byte row stores.
//                              staged through shared memory so that all global traffic is row-contiguous/coalesced.
//                              With split-K the partial tile goes to an L2-resident workspace and a second, fully
//                              parallel kernel sums the splits in fixed order (deterministic) and runs the epilogue.
//
// Reference call sites are the same as gemm_simt.cu (which remains the exact-fp32 referee and the fallback for
// shapes this kernel does not take: K % 32 != 0, N < 64, strided / upsampling convs).
#include <cuda.h>

#include "common.cuh"

namespace mugd {

#ifndef MUGD_TC_DECOUPLED
#define MUGD_TC_DECOUPLED 1      // measured: Beff=64 step 14.43 -> 13.90 ms, conv3 640->256 k-step 1.10 -> 1.02 us (0 = coupled stages)
#endif
constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // fp32 elements per k-step = 128 bytes = one swizzle row
constexpr int TC_THREADS = 256;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB

struct TcParams {
    mugd_gemm g;
    float* ws;                // split-K partial tiles [tile][split][128][BN]
    int32_t* counters;        // one ticket per output tile, zero at rest
    int32_t splits;
    int32_t total_it;         // taps * K / 32
    int32_t kblocks;          // K / 32
    int32_t Lrows, Bs;        // row structure of the A tensor map (Lrows = rows per sample, Bs samples)
    int32_t box_l, box_b;     // TMA box: box_l rows of box_b consecutive samples (box_l*box_b <= 128)
    int32_t tiles_per_sample; // when Lrows >= 128
    int32_t single_pass;      // 1: plain TF32 (a_hi*w_hi only, ~2^-11 relative) -- opt-in speed mode, NOT used for parity/bench
    int32_t inkernel_reduce;  // split-K: the last-arriving CTA of a tile reduces it (few splits), no second launch
    int32_t cluster;          // split-K CTAs of a tile form a thread-block cluster and reduce through DSMEM
    int32_t dbg_plain_store;  // measurement aid: bare store loop for epilogue-free GEMMs
    int32_t pdl_reduce;       // split-K: the reduce kernel is a programmatic dependent launch (resident and waiting while the GEMM runs)
    long long* dbg;           // optional: CTA (0,0,0) writes globaltimer stamps {entry, setup done, accumulator ready, tile staged, epilogue done}
};

// ---- raw PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
    return t;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug traps (CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from tensor memory (lane = row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
// multicast variants for a cluster of MC CTAs along M that share the B (weight) tiles: every CTA loads 1/MC of the tile and
// the TMA writes it -- and signals the mbarrier at the same offset -- in all MC CTAs; the stage is released with a commit
// that arrives on every CTA's "empty" barrier.  L2 -> SM traffic for B drops by MC.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// One lane of a converged warp.  tcgen05.mma / tcgen05.commit / cp.async.bulk.tensor are uniform-datapath instructions: issued
// from a lane-divergent branch (`if (lane == 0)`) ptxas wraps every one of them in an elect-and-branch loop (~95 cycles per
// MMA measured, which starved the tensor pipe); guarded by elect.sync in a converged warp they issue back to back.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between 8-row
// groups | version=1 [46,48) | layout_type=SWIZZLE_128B(2) [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int BN, bool AT>
struct TcSmem {
    static constexpr uint32_t B_BYTES = BN * TC_BK * 4;
    static constexpr uint32_t A_BUFS = AT ? 1 : 2;            // AT: only the raw tile lives in smem (hi/lo go to TMEM)
    static constexpr uint32_t STAGE_BYTES = A_BUFS * TC_A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = AT ? ((BN == 256) ? 2 : (BN == 128 ? 4 : 6)) : ((BN == 256) ? 2 : (BN == 128 ? 3 : 4));
    // Decoupled rings (256-wide tiles): only two 64 KB weight stages fit, and tied to the A tile they sat idle while the
    // activations were fetched and split.  Decoupled, the A side is a 2-deep smem ring feeding a 4-deep ring of TMEM operand
    // slots and runs ahead, and the freed shared memory holds a THIRD weight stage; a weight stage is occupied only from its TMA
    // to the retirement of its MMAs.
    static constexpr bool DEC = AT && BN == 256 && (MUGD_TC_DECOUPLED != 0);
    static constexpr int SAS = DEC ? 2 : STAGES;           // raw activation tiles in shared memory
    static constexpr int SA = DEC ? 4 : STAGES;            // split activation tiles in tensor memory
    static constexpr int SW = DEC ? 3 : STAGES;            // weight stages (hi + lo)
    static constexpr uint32_t TILE_BYTES = DEC ? SAS * TC_A_BYTES + SW * 2 * B_BYTES : STAGES * STAGE_BYTES;
    static constexpr uint32_t TOTAL = TILE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Fused epilogue math on 4 consecutive accumulator columns.  ACT / GATE are compile-time so that the compiler
// cannot if-convert the branches into "compute SiLU, GELU and both gates for every element, then select"
// (which it did, costing ~4 us per tile); callers dispatch once per kernel on the (uniform) act/gate values.
template <int ACT, int GATE>
__device__ __forceinline__ void tc_finish4(const mugd_gemm& g, float4 acc, float4 bia, float4 rvv, float4 res, int m, int nn) {
    float x[4] = {acc.x + bia.x + rvv.x, acc.y + bia.y + rvv.y, acc.z + bia.z + rvv.z, acc.w + bia.w + rvv.w};
    if constexpr (ACT == MUGD_ACT_SILU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = silu_f(x[j]);
    } else if constexpr (ACT == MUGD_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = gelu_f(x[j]);
    }
    if constexpr (GATE == MUGD_GATE_NONE) {
        st_f4(g.C + (int64_t)m * g.ldc + nn, make_float4(x[0] + res.x, x[1] + res.y, x[2] + res.z, x[3] + res.w));
    } else {
        float o0, o1;
        if constexpr (GATE == MUGD_GATE_GEGLU) { o0 = x[0] * gelu_f(x[1]); o1 = x[2] * gelu_f(x[3]); }
        else { o0 = x[0] * sigmoid_f(x[1]); o1 = x[2] * sigmoid_f(x[3]); }
        const int no = nn >> 1;
        if (g.residual) {
            const float2 rr = *reinterpret_cast<const float2*>(g.residual + (int64_t)m * g.ldr + no);
            o0 += rr.x; o1 += rr.y;
        }
        *reinterpret_cast<float2*>(g.C + (int64_t)m * g.ldc + no) = make_float2(o0, o1);
    }
}

// phase 2 of the epilogue for one CTA: read the staged accumulator tile from shared memory (row pitch BN+4) and
// finish it with coalesced global traffic; U float4 per thread in flight, every global load issued before any use.
template <int BN, int ACT, int GATE>
__device__ __forceinline__ void tc_store_tile(const mugd_gemm& g, uint32_t stage, int m_base, int n0, int rows_valid, const float* rowvec) {
    constexpr int SP = BN + 4;
    constexpr int C4 = BN / 4;
    constexpr int U = 8;
#pragma unroll 1
    for (int i0 = 0; i0 < TC_BM * C4; i0 += TC_THREADS * U) {
        float4 acc[U], bia[U], rvv[U], res[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = i0 + u * TC_THREADS + (int)threadIdx.x;
            const int row = idx / C4, c4 = idx - row * C4;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(acc[u].x), "=f"(acc[u].y), "=f"(acc[u].z), "=f"(acc[u].w)
                         : "r"(stage + (uint32_t)(row * SP + c4 * 4) * 4u));
            const int m = m_base + row, nn = n0 + c4 * 4;
            ok[u] = row < rows_valid && m < g.M && nn < g.N;
            bia[u] = rvv[u] = res[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[u]) {
                if (g.bias) bia[u] = ld_f4(g.bias + nn);
                if (rowvec) rvv[u] = ld_f4(rowvec + (int64_t)(m / g.Lout) * g.rowvec_b_stride + nn);
                if (GATE == MUGD_GATE_NONE && g.residual) res[u] = ld_f4(g.residual + (int64_t)m * g.ldr + nn);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const int idx = i0 + u * TC_THREADS + (int)threadIdx.x;
            const int row = idx / C4, c4 = idx - row * C4;
            tc_finish4<ACT, GATE>(g, acc[u], bia[u], rvv[u], res[u], m_base + row, n0 + c4 * 4);
        }
    }
}

// single float4 variant used by the split-K reduce kernel
template <int ACT, int GATE>
__device__ __forceinline__ void tc_epi4(const mugd_gemm& g, float4 acc, int m, int nn, const float* rowvec) {
    float4 bia = make_float4(0.f, 0.f, 0.f, 0.f), rvv = bia, res = bia;
    if (g.bias) bia = ld_f4(g.bias + nn);
    if (rowvec) rvv = ld_f4(rowvec + (int64_t)(m / g.Lout) * g.rowvec_b_stride + nn);
    if (GATE == MUGD_GATE_NONE && g.residual) res = ld_f4(g.residual + (int64_t)m * g.ldr + nn);
    tc_finish4<ACT, GATE>(g, acc, bia, rvv, res, m, nn);
}

#define TC_DISPATCH_EPI(g, CALL)                                                                      \
    do {                                                                                              \
        if ((g).gate == MUGD_GATE_GEGLU) { CALL(MUGD_ACT_NONE, MUGD_GATE_GEGLU); }                    \
        else if ((g).gate == MUGD_GATE_GLU) { CALL(MUGD_ACT_NONE, MUGD_GATE_GLU); }                   \
        else if ((g).act == MUGD_ACT_SILU) { CALL(MUGD_ACT_SILU, MUGD_GATE_NONE); }                   \
        else if ((g).act == MUGD_ACT_GELU) { CALL(MUGD_ACT_GELU, MUGD_GATE_NONE); }                   \
        else { CALL(MUGD_ACT_NONE, MUGD_GATE_NONE); }                                                 \
    } while (0)

// AT = true: the converter writes a_hi / a_lo straight into tensor memory (tcgen05.st) and the MMAs take A from TMEM
// (.kind::tf32 "TS" form).  That removes the converter's 32 KB of shared-memory writes and the 3 x 16 KB of A-operand
// reads per k-step from the shared-memory port, which is what bounds the all-smem ("SS") variant.
// MC > 1: launched as clusters (1, MC, 1) of MC vertically adjacent output tiles that share their weight tiles via TMA multicast.
template <int BN, bool AT, int MC>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmWhi,
               const __grid_constant__ CUtensorMap tmWlo, const TcParams p) {
    using S = TcSmem<BN, AT>;
    constexpr int STAGES = S::STAGES;
    // TMEM columns: accumulator [0, BN), then (AT) per stage 32 columns a_hi + 32 columns a_lo
    constexpr bool DEC = S::DEC && MC == 1;
    constexpr int SAS = DEC ? S::SAS : STAGES;          // raw A tiles in smem
    constexpr int SA = DEC ? S::SA : STAGES;            // TMEM operand slots
    constexpr int SW = DEC ? S::SW : STAGES;            // weight ring
    constexpr int TMEM_NEED = AT ? BN + SA * 64 : BN;
    constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512));
    static_assert(TMEM_NEED <= 512, "tensor memory budget");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B needs 1024-B alignment
    const uint32_t bars = base + S::TILE_BYTES;                          // barrier block (8-byte aligned)
    // barrier addresses: full[s], conv[s], empty[s], accum ; tmem ptr slot after them
    // decoupled rings add: afree[s] (A smem slot read by the converter), wfull[s] / wfree[s] (weight stage landed / retired);
    // bar_empty[s] then means "TMEM operand slot s retired"
    auto bar_full = [&](int s) { return bars + 8u * s; };                              // [SAS] raw A tile landed
    auto bar_conv = [&](int s) { return bars + 8u * (SAS + s); };                      // [SA]  split A in its TMEM slot
    auto bar_empty = [&](int s) { return bars + 8u * (SAS + SA + s); };                // [SA]  coupled: stage free; decoupled: TMEM slot retired
    auto bar_afree = [&](int s) { return bars + 8u * (SAS + 2 * SA + s); };            // [SAS] raw A tile consumed
    auto bar_wfull = [&](int s) { return bars + 8u * (2 * SAS + 2 * SA + s); };        // [SW]
    auto bar_wfree = [&](int s) { return bars + 8u * (2 * SAS + 2 * SA + SW + s); };   // [SW]
    const uint32_t bar_accum = bars + 8u * (DEC ? 2 * SAS + 2 * SA + 2 * SW : 3 * STAGES);
    const uint32_t tmem_slot = bar_accum + 8u;
    static_assert(8 * (DEC ? 2 * SAS + 2 * SA + 2 * SW + 2 : 3 * STAGES + 2) <= 256, "barrier block");
    auto a_hi = [&](int s) { return DEC ? base + s * TC_A_BYTES : base + s * S::STAGE_BYTES; };
    auto a_lo = [&](int s) { return base + s * S::STAGE_BYTES + TC_A_BYTES; };
    auto b_hi = [&](int s) { return DEC ? base + SAS * TC_A_BYTES + s * 2 * S::B_BYTES : base + s * S::STAGE_BYTES + S::A_BUFS * TC_A_BYTES; };
    auto b_lo = [&](int s) { return b_hi(s) + S::B_BYTES; };

    const mugd_gemm& g = p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN;
    // ---- tile -> rows -------------------------------------------------------------------------------
    int b_base, l_base, rows_valid;
    if (p.Lrows >= TC_BM) {
        b_base = blockIdx.y / p.tiles_per_sample;
        l_base = (blockIdx.y % p.tiles_per_sample) * TC_BM;
        rows_valid = min(TC_BM, p.Lrows - l_base);
    } else {
        b_base = blockIdx.y * p.box_b;
        l_base = 0;
        rows_valid = min(p.box_b, p.Bs - b_base) * p.Lrows;
    }
    const int m_base = b_base * p.Lrows + l_base;
    const int it_begin = (int)(((long long)p.total_it * blockIdx.z) / p.splits);
    const int it_end = (int)(((long long)p.total_it * (blockIdx.z + 1)) / p.splits);
    const int nit = it_end - it_begin;

    pdl_trigger_gemm_entry();           // (experiment switch, off: see common.cuh)
    if (p.pdl_reduce) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the split-K reduce kernel may take its seats now
    const uint32_t crank = (MC > 1) ? cluster_rank() : 0u;
    const bool dbg_cta = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    // ---- one-time setup ------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int s = 0; s < SA; ++s) {
            mbar_init(bar_conv(s), 4);        // one arrival per converter warp
            mbar_init(bar_empty(s), MC);      // one commit per CTA of the cluster
        }
        for (int s = 0; s < SAS; ++s) {
            mbar_init(bar_full(s), 1);
            if constexpr (DEC) mbar_init(bar_afree(s), 4);
        }
        if constexpr (DEC) {
            for (int s = 0; s < SW; ++s) {
                mbar_init(bar_wfull(s), 1);
                mbar_init(bar_wfree(s), 1);
            }
        }
        mbar_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (MC > 1) cluster_sync_all();   // peers' barriers must exist before any multicast can signal them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // barriers, TMEM and descriptors are set up: from here on global memory written by the previous kernel is touched
    pdl_wait();
    if (dbg_cta && threadIdx.x == 0) { p.dbg[0] = gtimer(); p.dbg[1] = p.dbg[0]; }

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        // the whole warp walks the loop converged; one elected lane issues the copies
        {
            const uint32_t a_tx = (uint32_t)(p.box_l * p.box_b) * TC_BK * 4;
            for (int i = 0; i < nit; ++i) {
                const int s = i % SAS;
                const uint32_t ph = (uint32_t)(i / SAS) & 1u;
                if constexpr (DEC) mbar_wait(bar_afree(s), ph ^ 1u);
                else mbar_wait(bar_empty(s), ph ^ 1u);
                if (elect_one()) {
                if (dbg_cta && i < 24) p.dbg[8 + i * 6 + 5] = gtimer();
                const int it = it_begin + i;
                const int t = it / p.kblocks;
                const int kb = it - t * p.kblocks;
                mbar_expect_tx(bar_full(s), DEC ? a_tx : a_tx + (p.single_pass ? 1u : 2u) * S::B_BYTES);
                // row addressing per tap: SAME = l+t-1, TAPS = l+t+shift (zero fill outside the sample by TMA bounds);
                // DOWN (stride 2, right pad) uses one strided tensor map per tap (row l of map t = source row 2l+t)
                const CUtensorMap* ma = &tmA;
                int lshift = 0;
                if (g.conv_mode == MUGD_CONV_SAME) lshift = t - 1;
                else if (g.conv_mode == MUGD_CONV_TAPS) lshift = (t + g.tap_shift) * (g.tap_dilation > 1 ? g.tap_dilation : 1);
                else if (g.conv_mode == MUGD_CONV_DOWN) ma = (t == 0) ? &tmA : (t == 1 ? &tmA1 : &tmA2);
                tma_load_3d(a_hi(s), ma, bar_full(s), kb * TC_BK, l_base + lshift, b_base);
                if constexpr (MC > 1) {
                    const int part = BN / MC;                       // this CTA's share of the weight tile rows
                    const uint32_t doff = crank * (uint32_t)part * (TC_BK * 4);
                    tma_load_2d_mc(b_hi(s) + doff, &tmWhi, bar_full(s), t * g.K + kb * TC_BK, n0 + (int)crank * part, (uint16_t)((1u << MC) - 1u));
                    if (!p.single_pass)
                        tma_load_2d_mc(b_lo(s) + doff, &tmWlo, bar_full(s), t * g.K + kb * TC_BK, n0 + (int)crank * part, (uint16_t)((1u << MC) - 1u));
                } else if constexpr (!DEC) {
                    tma_load_2d(b_hi(s), &tmWhi, bar_full(s), t * g.K + kb * TC_BK, n0);
                    if (!p.single_pass) tma_load_2d(b_lo(s), &tmWlo, bar_full(s), t * g.K + kb * TC_BK, n0);
                }
                if (dbg_cta && i < 24) p.dbg[8 + i * 6 + 0] = gtimer();
                }
                __syncwarp();
            }
        }
    } else if (DEC && warp == 3) {
        // ===================================== weight producer (decoupled rings) ================
        if constexpr (DEC) {
            for (int i = 0; i < nit; ++i) {
                const int s = i % SW;
                const uint32_t ph = (uint32_t)(i / SW) & 1u;
                mbar_wait(bar_wfree(s), ph ^ 1u);
                if (elect_one()) {
                    const int it = it_begin + i;
                    const int t = it / p.kblocks;
                    const int kb = it - t * p.kblocks;
                    mbar_expect_tx(bar_wfull(s), (p.single_pass ? 1u : 2u) * S::B_BYTES);
                    tma_load_2d(b_hi(s), &tmWhi, bar_wfull(s), t * g.K + kb * TC_BK, n0);
                    if (!p.single_pass) tma_load_2d(b_lo(s), &tmWlo, bar_wfull(s), t * g.K + kb * TC_BK, n0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2,
            // A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int i = 0; i < nit; ++i) {
                const int s = i % SA;
                const uint32_t ph = (uint32_t)(i / SA) & 1u;
                const int sw = DEC ? i % SW : s;
                mbar_wait(bar_conv(s), ph);
                if constexpr (DEC) mbar_wait(bar_wfull(sw), (uint32_t)(i / SW) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                if (dbg_cta && i < 24) p.dbg[8 + i * 6 + 3] = gtimer();
                const uint64_t dbh = umma_desc(b_hi(sw)), dbl = umma_desc(b_lo(sw));
                if constexpr (AT) {
                    const uint32_t ta_hi = tmem_base + (uint32_t)(BN + s * 64), ta_lo = ta_hi + 32u;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk) {
                        const uint64_t ko = (uint64_t)(kk * 2);     // 8 fp32 = 32 bytes = 2 x 16-byte units
                        if (p.single_pass) {
                            umma_tf32_ts(tmem_base, ta_hi + kk * 8, dbh + ko, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                        } else {
                            umma_tf32_ts(tmem_base, ta_lo + kk * 8, dbh + ko, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                            umma_tf32_ts(tmem_base, ta_hi + kk * 8, dbl + ko, idesc, 1u);
                            umma_tf32_ts(tmem_base, ta_hi + kk * 8, dbh + ko, idesc, 1u);
                        }
                    }
                } else {
                    const uint64_t dah = umma_desc(a_hi(s)), dal = umma_desc(a_lo(s));
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk) {
                        const uint64_t ko = (uint64_t)(kk * 2);
                        if (p.single_pass) {
                            umma_tf32(tmem_base, dah + ko, dbh + ko, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                        } else {
                            umma_tf32(tmem_base, dal + ko, dbh + ko, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                            umma_tf32(tmem_base, dah + ko, dbl + ko, idesc, 1u);
                            umma_tf32(tmem_base, dah + ko, dbh + ko, idesc, 1u);
                        }
                    }
                }
                if constexpr (MC > 1) umma_commit_mc(bar_empty(s), (uint16_t)((1u << MC) - 1u));
                else umma_commit(bar_empty(s));                       // stage (decoupled: TMEM operand slot) reusable once these MMAs retire
                if constexpr (DEC) umma_commit(bar_wfree(sw));        // ... and the weight stage
                if (dbg_cta && i < 24) p.dbg[8 + i * 6 + 4] = gtimer();
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(bar_accum);
        }
    } else if (warp >= 4) {
        // ===================================== converter ========================================
        const int ct = threadIdx.x - 128;                             // 0..127
        for (int i = 0; i < nit; ++i) {
            const int s = i % SA;                                     // TMEM operand slot
            const int sm = i % SAS;                                   // raw tile in shared memory
            mbar_wait(bar_full(sm), (uint32_t)(i / SAS) & 1u);
            if constexpr (DEC) {
                mbar_wait(bar_empty(s), ((uint32_t)(i / SA) & 1u) ^ 1u);   // the MMAs that read TMEM slot s last time have retired
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            if (dbg_cta && ct == 0 && i < 24) p.dbg[8 + i * 6 + 1] = gtimer();
            if constexpr (AT) {
                // thread = tile row (= TMEM lane): read the row's 128 bytes out of the 128B-swizzled tile (16-byte chunk c
                // of row r sits at chunk c ^ (r & 7)), split, and store hi / lo to this stage's TMEM columns
                const int r = (warp & 3) * 32 + lane;
                const uint32_t rowaddr = a_hi(sm) + (uint32_t)r * 128u;
                float hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 x;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                                 : "r"(rowaddr + (uint32_t)((c ^ (r & 7)) * 16)));
                    hi[c * 4] = to_tf32(x.x); hi[c * 4 + 1] = to_tf32(x.y); hi[c * 4 + 2] = to_tf32(x.z); hi[c * 4 + 3] = to_tf32(x.w);
                    lo[c * 4] = to_tf32(x.x - hi[c * 4]); lo[c * 4 + 1] = to_tf32(x.y - hi[c * 4 + 1]);
                    lo[c * 4 + 2] = to_tf32(x.z - hi[c * 4 + 2]); lo[c * 4 + 3] = to_tf32(x.w - hi[c * 4 + 3]);
                }
                const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(BN + s * 64);
                tmem_st32(ta, hi);
                tmem_st32(ta + 32u, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            } else {
            const uint32_t hi_addr = a_hi(s), lo_addr = a_lo(s);
#pragma unroll
            for (int j = 0; j < (int)(TC_A_BYTES / 16 / 128); ++j) {   // 8 x 16 bytes per thread
                const uint32_t off = (uint32_t)(ct + j * 128) * 16u;
                float4 x;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(hi_addr + off));
                float4 h, l;
                h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
                l.x = to_tf32(x.x - h.x); l.y = to_tf32(x.y - h.y); l.z = to_tf32(x.z - h.z); l.w = to_tf32(x.w - h.w);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hi_addr + off), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr + off), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_conv(s));
                if constexpr (DEC) mbar_arrive(bar_afree(sm));        // the raw tile has been read: its smem slot may be refilled
            }
            if (dbg_cta && ct == 0 && i < 24) p.dbg[8 + i * 6 + 2] = gtimer();
        }
        // ===================================== epilogue =========================================
        mbar_wait(bar_accum, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg_cta && ct == 0) p.dbg[2] = gtimer();
        const int q = warp & 3;                                        // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;                                   // tile row == TMEM lane
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        // phase 1: TMEM -> registers -> shared (the pipeline buffers are free: every TMA landed, every MMA retired).
        // Row pitch BN+4 floats keeps the per-row float4 stores and the row-contiguous reads below conflict-free.
        constexpr int SP = BN + 4;
        const uint32_t stage = base;
        float v[32];
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            tmem_ld32(trow + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)(r * SP + c0 + j * 4) * 4u), "f"(v[j * 4]),
                             "f"(v[j * 4 + 1]), "f"(v[j * 4 + 2]), "f"(v[j * 4 + 3]) : "memory");
        }
        if (dbg_cta && ct == 0) p.dbg[3] = gtimer();
    }
    // ---- phase 2 (all 8 warps): consecutive threads take consecutive float4 of a row -> coalesced global traffic.
    // The accumulator tile is complete in shared memory once the 4 epilogue warps pass this barrier; the producer /
    // MMA / allocator warps have nothing left to do and join in.
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    pdl_trigger_late();                 // (experiment switch, off) only the store phase is left
    if (dbg_cta && threadIdx.x == 0) p.dbg[5] = gtimer();
    {
        const int step = g.step ? *g.step : 0;
        const float* rowvec = g.rowvec ? g.rowvec + (int64_t)step * g.rowvec_step_stride : nullptr;
        const uint32_t stage = base;
        if (p.splits > 1 && p.cluster) {
            // ---- split-K reduction through distributed shared memory: the `splits` CTAs of this tile are one cluster
            // (1,1,splits).  Every CTA owns a band of rows, sums that band over all peers' staged tiles in rank order
            // (deterministic), finishes it with the fused epilogue.  No workspace round trip, no second launch.
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
            constexpr int SP = BN + 4;
            constexpr int C4 = BN / 4;
            const int rows_per = (TC_BM + p.splits - 1) / p.splits;
            const int r0 = (int)blockIdx.z * rows_per;
            const int r1 = min(TC_BM, r0 + rows_per);
            const int n4 = max(0, r1 - r0) * C4;
            for (int i = (int)threadIdx.x; i < n4; i += TC_THREADS) {
                const int row = r0 + i / C4, c4 = i % C4;
                const uint32_t laddr = stage + (uint32_t)(row * SP + c4 * 4) * 4u;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int z = 0; z < p.splits; ++z) {
                    uint32_t raddr;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(z));
                    float4 t4;
                    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t4.x), "=f"(t4.y), "=f"(t4.z), "=f"(t4.w) : "r"(raddr));
                    acc.x += t4.x; acc.y += t4.y; acc.z += t4.z; acc.w += t4.w;
                }
                const int m = m_base + row, nn = n0 + c4 * 4;
                if (row < rows_valid && m < g.M && nn < g.N) {
#define TC_CALL_EPI(A_, G_) tc_epi4<A_, G_>(g, acc, m, nn, rowvec)
                    TC_DISPATCH_EPI(g, TC_CALL_EPI);
#undef TC_CALL_EPI
                }
            }
            // peers may still be reading this CTA's tile: leave together
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else if (p.splits > 1) {
            const int tile_lin = blockIdx.y * gridDim.x + blockIdx.x;
            float* wsp = p.ws + ((int64_t)tile_lin * p.splits + blockIdx.z) * (TC_BM * BN);
            constexpr int SP = BN + 4;
            constexpr int C4 = BN / 4;
            constexpr int U = 8;
#pragma unroll 1
            for (int i0 = 0; i0 < TC_BM * C4; i0 += TC_THREADS * U) {
                float4 acc[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int idx = i0 + u * TC_THREADS + (int)threadIdx.x;
                    const int row = idx / C4, c4 = idx - row * C4;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(acc[u].x), "=f"(acc[u].y), "=f"(acc[u].z), "=f"(acc[u].w)
                                 : "r"(stage + (uint32_t)(row * SP + c4 * 4) * 4u));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) st_f4(wsp + (i0 + u * TC_THREADS + (int)threadIdx.x) * 4, acc[u]);   // [row][BN] dense
            }
            if (p.inkernel_reduce == 2) {
                // Cooperative launch (all CTAs of the grid are co-resident, guaranteed by the driver): the `splits` CTAs of a
                // tile meet at the tile's counter once their partial tiles are in L2, then each sums ITS band of rows over all
                // partials in fixed split order (deterministic) and finishes it with the fused epilogue.  The reduction is
                // spread over the same CTAs that produced it: no second launch and no serial tail.
                int* cnt = p.counters + tile_lin;
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) {
                    atomicAdd(cnt, 1);
                    const long long t0 = clock64();
                    int seen;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
                        if (clock64() - t0 > 4000000000LL) __trap();      // a protocol bug traps instead of hanging the GPU
                    } while (seen < p.splits);
                }
                __syncthreads();
                __threadfence();
                const int rows_per = (TC_BM + p.splits - 1) / p.splits;
                const int r0 = (int)blockIdx.z * rows_per;
                const int r1 = min(TC_BM, r0 + rows_per);
                const int n4 = max(0, r1 - r0) * C4;
                const float* wst = p.ws + ((int64_t)tile_lin * p.splits) * (TC_BM * BN);
                for (int i = (int)threadIdx.x; i < n4; i += TC_THREADS) {
                    const int row = r0 + i / C4, c4 = i % C4;
                    const float* src = wst + (int64_t)row * BN + c4 * 4;
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                    for (int z = 0; z < p.splits; ++z) {
                        const float4 t4 = __ldcg(reinterpret_cast<const float4*>(src + (int64_t)z * (TC_BM * BN)));
                        sum.x += t4.x; sum.y += t4.y; sum.z += t4.z; sum.w += t4.w;
                    }
                    const int m = m_base + row, nn = n0 + c4 * 4;
                    if (row < rows_valid && m < g.M && nn < g.N) {
#define TC_CALL_EPI(A_, G_) tc_epi4<A_, G_>(g, sum, m, nn, rowvec)
                        TC_DISPATCH_EPI(g, TC_CALL_EPI);
#undef TC_CALL_EPI
                    }
                }
                __syncthreads();
                // second round of tickets: the CTA that completes it puts the counter back to rest (nobody can still be
                // polling: every CTA left the wait above before taking its second ticket)
                if (threadIdx.x == 0 && atomicAdd(cnt, 1) == 2 * p.splits - 1) atomicExch(cnt, 0);
            } else if (p.inkernel_reduce) {
                // Few splits: the CTA that arrives last at the tile's ticket sums all partial tiles (fixed split order ->
                // deterministic) straight out of L2 and runs the epilogue; nobody waits, so there is no co-residency
                // requirement, and the second launch is saved.  Many splits keep the fully parallel reduce kernel.
                __shared__ int s_ticket;
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) s_ticket = atomicAdd(p.counters + tile_lin, 1);
                __syncthreads();
                if (s_ticket == p.splits - 1) {
                    __threadfence();
                    const float* wst = p.ws + ((int64_t)tile_lin * p.splits) * (TC_BM * BN);
#pragma unroll 1
                    for (int i0 = 0; i0 < TC_BM * C4; i0 += TC_THREADS * 4) {
                        float4 sum[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) sum[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int z = 0; z < p.splits; ++z) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float4 t4 = __ldcg(reinterpret_cast<const float4*>(
                                    wst + (int64_t)z * (TC_BM * BN) + (int64_t)(i0 + u * TC_THREADS + (int)threadIdx.x) * 4));
                                sum[u].x += t4.x; sum[u].y += t4.y; sum[u].z += t4.z; sum[u].w += t4.w;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int idx = i0 + u * TC_THREADS + (int)threadIdx.x;
                            const int row = idx / C4, c4 = idx - row * C4;
                            const int m = m_base + row, nn = n0 + c4 * 4;
                            if (row < rows_valid && m < g.M && nn < g.N) {
#define TC_CALL_EPI(A_, G_) tc_epi4<A_, G_>(g, sum[u], m, nn, rowvec)
                                TC_DISPATCH_EPI(g, TC_CALL_EPI);
#undef TC_CALL_EPI
                            }
                        }
                    }
                    if (threadIdx.x == 0) p.counters[tile_lin] = 0;          // ticket back to rest for the next launch / replay
                }
            }
        } else if (p.dbg_plain_store && !g.bias && !rowvec && !g.residual && g.act == MUGD_ACT_NONE && g.gate == MUGD_GATE_NONE) {
            // measurement aid (mugd_debug_set_tc_plain_store): the split-K path's bare store loop writing C rows -- 1.4 us per
            // 128x128 tile against 3.1 us for tc_store_tile with nothing to add
            constexpr int SP = BN + 4;
            constexpr int C4 = BN / 4;
            constexpr int U = 8;
#pragma unroll 1
            for (int i0 = 0; i0 < TC_BM * C4; i0 += TC_THREADS * U) {
                float4 acc[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int idx = i0 + u * TC_THREADS + (int)threadIdx.x;
                    const int row = idx / C4, c4 = idx - row * C4;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(acc[u].x), "=f"(acc[u].y), "=f"(acc[u].z), "=f"(acc[u].w)
                                 : "r"(stage + (uint32_t)(row * SP + c4 * 4) * 4u));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int idx = i0 + u * TC_THREADS + (int)threadIdx.x;
                    const int row = idx / C4, c4 = idx - row * C4;
                    const int m = m_base + row, nn = n0 + c4 * 4;
                    if (row < rows_valid && m < g.M && nn < g.N) st_f4(g.C + (int64_t)m * g.ldc + nn, acc[u]);
                }
            }
        } else {
#define TC_CALL_STORE(A_, G_) tc_store_tile<BN, A_, G_>(g, stage, m_base, n0, rows_valid, rowvec)
            TC_DISPATCH_EPI(g, TC_CALL_STORE);
#undef TC_CALL_STORE
        }
        if (dbg_cta && threadIdx.x == 0) p.dbg[4] = gtimer();
    }
    // ---- teardown (all tcgen05.ld completed before the phase-2 barrier) ----------------------------------
    if constexpr (MC > 1) cluster_sync_all();   // trailing multicast commits must not land in an exited CTA
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// split-K second pass: sum the partial tiles in fixed split order (deterministic) and run the fused epilogue.
// One thread per (row, 32-column chunk); fully parallel over the GPU and L2-resident.
template <int BN>
__global__ void __launch_bounds__(256)
gemm_tc_reduce_kernel(const TcParams p, int gx, int gy) {
    pdl_trigger();
    pdl_wait();
    const mugd_gemm& g = p.g;
    constexpr int C4 = BN / 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)gx * gy * TC_BM * C4;
    if (idx >= total) return;
    const int c4 = (int)(idx % C4);
    const int r = (int)((idx / C4) % TC_BM);
    const int tile_lin = (int)(idx / ((long long)C4 * TC_BM));
    const int bx = tile_lin % gx, by = tile_lin / gx;
    int b_base, l_base, rows_valid;
    if (p.Lrows >= TC_BM) {
        b_base = by / p.tiles_per_sample;
        l_base = (by % p.tiles_per_sample) * TC_BM;
        rows_valid = min(TC_BM, p.Lrows - l_base);
    } else {
        b_base = by * p.box_b;
        l_base = 0;
        rows_valid = min(p.box_b, p.Bs - b_base) * p.Lrows;
    }
    const int m = b_base * p.Lrows + l_base + r;
    const int n = bx * BN + c4 * 4;
    if (r >= rows_valid || m >= g.M || n >= g.N) return;
    const float* src = p.ws + ((long long)tile_lin * p.splits) * (TC_BM * BN) + (long long)r * BN + c4 * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int z = 0; z < p.splits; ++z) {                                   // fixed order -> deterministic
        const float4 t4 = __ldcg(reinterpret_cast<const float4*>(src + (long long)z * (TC_BM * BN)));
        acc.x += t4.x; acc.y += t4.y; acc.z += t4.z; acc.w += t4.w;
    }
    const int step = g.step ? *g.step : 0;
    const float* rowvec = g.rowvec ? g.rowvec + (int64_t)step * g.rowvec_step_stride : nullptr;
#define TC_CALL_EPI(A_, G_) tc_epi4<A_, G_>(g, acc, m, n, rowvec)
    TC_DISPATCH_EPI(g, TC_CALL_EPI);
#undef TC_CALL_EPI
}

static long long* g_tc_dbg = nullptr;
static bool g_tc_a_in_tmem = true;   // A operand of the MMAs from tensor memory (TS form) instead of shared memory (SS)
// Split counts up to this value reduce inside the GEMM kernel (last-arriving CTA of a tile, no second launch).  Measured
// SLOWER on B200 (GEMM family 4.17 ms vs 3.01 ms per step at 4; worse at 8/16): one CTA pulling splits x 64 KB out of L2
// costs more than the ~3 us reduce launch -> 0 (off) by default.
static int g_tc_inkernel_max = 0;
static bool g_tc_single_pass = false; // opt-in plain-TF32 mode (one product instead of three)
// Weight-tile TMA multicast over clusters of 2/4 vertically adjacent tiles: measured on B200 it does not help (B=32 step:
// GEMM family 11.25 ms unicast, 11.57 ms clusters of 2, 11.74 ms clusters of 4).  The large-GEMM main loop is bound by
// the chip-wide L2 throughput (~42 B/clk/SM with all SMs pulling), but L2 already merges the requests of the few SMs that
// read the same weight line at the same time, so multicast at cluster sizes <= 4 saves no L2 bandwidth and only adds
// the cluster launch/sync cost -> off by default, kept for experiments (MUGD_TC_MC=2|4).
// 64-wide tiles for GEMMs that could use 128 (more, smaller CTAs for the grids that underfill the machine)
// planner constants, re-measured after the elect.sync issue fix (tools/bench_gemm.py, tools/gpu_cost.sh sweep):
// us per k-step of a 128- / 256-wide tile, us per split-K round trip (workspace + reduce launch)
static float g_tc_cost[3] = {0.55f, 0.9f, 4.0f};
static int g_tc_coop_reduce = 0;         // split-K: cooperative launch + per-tile rendezvous, reduction spread over the split CTAs
static float g_tc_coop_cost = 2.0f;      // planner: us per split round trip in that mode
static int g_tc_plain_store = 0;
static int g_tc_pdl_reduce = 0;
static int g_tc_narrow_tiles = 0;
static float g_tc_kstep64 = 0.4f;
static int g_tc_multicast = 0;        // max cluster size (along M) for weight-tile TMA multicast; 0/1 = off
static int g_tc_force_bn = 0;        // experiments: 0 = cost model, 128 / 256 = force the tile width where legal
// Split-K reduction through a thread-block cluster + DSMEM instead of workspace + reduce kernel.  Works (tests pass)
// but measured slower on B200: clusters of 197 KB-smem CTAs schedule poorly (8 co-resident SMs of one GPC) and DSMEM
// reads cost ~4 us per tile: GEMM family 4.50 ms vs 3.38 ms per step -> off by default, kept for experiments.
static bool g_tc_cluster = false;

// =====================================================================================================
// host side
// =====================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

struct TcGeometry {
    int BN, splits, gx, gy, Lrows, Bs, box_l, box_b, tiles_per_sample, total_it, mc;
    int64_t ws_floats;
};

bool gemm_tc_supported(const mugd_gemm& g) {
    if (!(g.conv_mode == MUGD_CONV_NONE || g.conv_mode == MUGD_CONV_SAME || g.conv_mode == MUGD_CONV_DOWN ||
          g.conv_mode == MUGD_CONV_TAPS)) return false;
    if (g.conv_mode == MUGD_CONV_DOWN && g.Lout < 2) return false;
    if (g.K % TC_BK != 0 || g.N < 64 || g.N % 4 != 0) return false;
    if (!g.W_hi || !g.W_lo) return false;
    if (g.lda % 4 != 0 || !aligned16(g.A) || !aligned16(g.W_hi) || !aligned16(g.W_lo)) return false;
    return true;
}

static TcGeometry tc_geometry(const mugd_gemm& g, int sm_count, int forced_split) {
    TcGeometry t;
    t.BN = (g.N >= 128) ? 128 : 64;
    if (g.conv_mode == MUGD_CONV_NONE) { t.Lrows = g.M; t.Bs = 1; }
    else { t.Lrows = g.Lout; t.Bs = g.M / g.Lout; }
    if (t.Lrows >= TC_BM) {
        t.box_l = TC_BM; t.box_b = 1;
        t.tiles_per_sample = (t.Lrows + TC_BM - 1) / TC_BM;
        t.gy = t.tiles_per_sample * t.Bs;
    } else {
        t.box_l = t.Lrows;
        t.box_b = TC_BM / t.Lrows;
        if (t.box_b > t.Bs) t.box_b = t.Bs;
        t.tiles_per_sample = 1;
        t.gy = (t.Bs + t.box_b - 1) / t.box_b;
    }
    t.total_it = g.taps * (g.K / TC_BK);
    // Cost model from the B200 micro-benchmark (tools/bench_gemm.py): a CTA needs ~1 us to fill its pipeline and
    // ~0.7 us per k-step with 128-wide tiles (~1.05 us with 256-wide tiles, which do twice the math per step but
    // only 2 pipeline stages fit); splitting K adds the workspace round trip and a second (reduce) launch, ~5 us.
    // Candidates: tile width 128 (or 64 for narrow N), 256 when N allows it, each with its best K split.
    int splits = 1;
    float best = 1e30f;
    const int tc_bn_env = g_tc_force_bn;
    static const int cands[3] = {64, 128, 256};
    for (int cand = 0; cand < 3; ++cand) {
        const int bn = cands[cand];
        if (bn > 64 && g.N < bn) break;
        if (bn == 64 && g.N >= 128 && !g_tc_narrow_tiles && tc_bn_env != 64) continue;
        if (tc_bn_env && bn != tc_bn_env && !(tc_bn_env > g.N && bn == (g.N >= 128 ? 128 : 64))) continue;
        const int gx = (g.N + bn - 1) / bn;
        const int tiles = gx * t.gy;
        const float kstep = bn == 256 ? g_tc_cost[1] : (bn == 128 ? g_tc_cost[0] : g_tc_kstep64);
        // 256-wide tiles only pay off unsplit (measured: l1/l2 FF1 and the B=64 convs gain 15-25 %, split cases lose)
        const int sp_max = forced_split > 0 ? forced_split : ((tiles < sm_count && bn != 256) ? (g_tc_cluster ? 8 : 16) : 1);
        for (int sp = forced_split > 0 ? forced_split : 1; sp <= sp_max && sp <= t.total_it; ++sp) {
            const int per = (t.total_it + sp - 1) / sp;
            if (forced_split <= 0 && sp > 1 && per < 2) break;
            if (forced_split <= 0 && sp > 1 && tiles * sp > 2 * sm_count) break;   // bounds the workspace: < 2*SMs partial tiles
            if (g_tc_coop_reduce && sp > 1 && tiles * sp > sm_count) break;          // cooperative reduce: the whole grid must be co-resident
            const int waves = (tiles * sp + sm_count - 1) / sm_count;
            // narrower tiles also shorten the epilogue (fewer columns per CTA): ~1 us per 64 columns on top of the fill
            const float fill = 1.0f + (g_tc_narrow_tiles ? 0.9f * (bn / 64 - 1) : 0.0f);
            const float est = waves * (fill + kstep * per) + (sp > 1 ? (g_tc_cluster ? 2.0f : (g_tc_coop_reduce ? g_tc_coop_cost : (sp <= g_tc_inkernel_max ? 1.0f + 0.6f * sp : g_tc_cost[2]))) : 0.0f);
            if (est < best - 0.25f) { best = est; splits = sp; t.BN = bn; }
        }
    }
    t.gx = (g.N + t.BN - 1) / t.BN;
    const int tiles = t.gx * t.gy;
    // weight-tile multicast: only worth it when the grid oversubscribes the machine (the main loop is then bound by the
    // L2 -> SM operand stream, ~39 B/clk/SM with all SMs pulling) and the tile rows pair up
    t.mc = 1;
    if (g_tc_multicast && g_tc_a_in_tmem && splits == 1 && t.BN >= 128 && tiles >= sm_count) {
        if (g_tc_multicast >= 4 && t.gy % 4 == 0) t.mc = 4;
        else if (t.gy % 2 == 0) t.mc = 2;
    }
    if (splits > t.total_it) splits = t.total_it;
    if (splits < 1) splits = 1;
    t.splits = splits;
    t.ws_floats = splits > 1 ? (int64_t)tiles * splits * TC_BM * t.BN : 0;
    return t;
}

template <int BN, bool AT, int MC>
static int tc_launch(const CUtensorMap* tmAs, const CUtensorMap& tmWhi, const CUtensorMap& tmWlo, const TcParams& p,
                     const TcGeometry& t, cudaStream_t st) {
    const CUtensorMap &tmA = tmAs[0], &tmA1 = tmAs[1], &tmA2 = tmAs[2];
    static bool configured = false;
    if (!configured) {
        MUGD_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, AT, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcSmem<BN, AT>::TOTAL));
        configured = true;
    }
    dim3 grid(t.gx, t.gy, t.splits);
    if (MC > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = TcSmem<BN, AT>::TOTAL;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1;
        attr[0].val.clusterDim.y = MC;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = g_use_pdl ? 2 : 1;
        MUGD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, AT, MC>, tmA, tmA1, tmA2, tmWhi, tmWlo, p));
        if (t.splits > 1 && !p.inkernel_reduce) {
            const long long total = (long long)t.gx * t.gy * TC_BM * (BN / 4);
            MUGD_CHECK_CUDA(launch_k(gemm_tc_reduce_kernel<BN>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, p, t.gx, t.gy));
        }
        return MUGD_OK;
    }
    if (p.cluster && t.splits > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = TcSmem<BN, AT>::TOTAL;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = (unsigned)t.splits;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = g_use_pdl ? 2 : 1;
        MUGD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, AT, MC>, tmA, tmA1, tmA2, tmWhi, tmWlo, p));
        return MUGD_OK;
    }
    if (p.inkernel_reduce == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = TcSmem<BN, AT>::TOTAL;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MUGD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, AT, MC>, tmA, tmA1, tmA2, tmWhi, tmWlo, p));
        return MUGD_OK;
    }
    MUGD_CHECK_CUDA(launch_k(gemm_tc_kernel<BN, AT, MC>, grid, dim3(TC_THREADS), TcSmem<BN, AT>::TOTAL, st, tmA, tmA1, tmA2, tmWhi, tmWlo, p));
    if (t.splits > 1 && !p.inkernel_reduce) {
        const long long total = (long long)t.gx * t.gy * TC_BM * (BN / 4);
        if (p.pdl_reduce) {
            // programmatic dependent launch: the reduce grid is scheduled while the GEMM still runs and sits in
            // griddepcontrol.wait until the GEMM grid has completed and flushed -> no launch gap between the two
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)((total + 255) / 256));
            cfg.blockDim = dim3(256);
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            MUGD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_reduce_kernel<BN>, p, t.gx, t.gy));
        } else {
            MUGD_CHECK_CUDA(launch_k(gemm_tc_reduce_kernel<BN>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, p, t.gx, t.gy));
        }
    }
    return MUGD_OK;
}

int launch_gemm_tc(const DeviceInfo& dev, const mugd_gemm& g, cudaStream_t st, int* launches) {
    MUGD_REQUIRE(gemm_tc_supported(g), "gemm_tc: unsupported shape/operands");
    EncodeTiledFn enc = get_encode();
    MUGD_REQUIRE(enc != nullptr, "gemm_tc: cuTensorMapEncodeTiled not available from the driver");
    const TcGeometry t = tc_geometry(g, dev.sm_count, g.split_k);
    const bool use_cluster = g_tc_cluster && t.splits > 1 && t.splits <= 8;
    if (t.splits > 1 && !use_cluster) {
        MUGD_REQUIRE(g.workspace, "gemm_tc: split-K needs a workspace");
        MUGD_REQUIRE(g.workspace_bytes >= t.ws_floats * 4, "gemm_tc: workspace too small (%lld < %lld)", (long long)g.workspace_bytes,
                     (long long)t.ws_floats * 4);
    }
    CUtensorMap tmAs[3], tmWhi, tmWlo;
    for (int tap = 0; tap < 3; ++tap) {
        if (tap > 0 && g.conv_mode != MUGD_CONV_DOWN) { tmAs[tap] = tmAs[0]; continue; }
        const bool down = g.conv_mode == MUGD_CONV_DOWN;
        // DOWN: row l of the map of tap t is source row 2l+t; the last row of tap 2 is the right padding -> out of bounds
        const cuuint64_t rows = down ? (cuuint64_t)(t.Lrows - (tap == 2 ? 1 : 0)) : (cuuint64_t)t.Lrows;
        const cuuint64_t sample_rows = down ? (cuuint64_t)g.Lin : (cuuint64_t)t.Lrows;
        cuuint64_t dims[3] = {(cuuint64_t)g.K, rows, (cuuint64_t)t.Bs};
        cuuint64_t strides[2] = {(cuuint64_t)g.lda * 4 * (down ? 2 : 1), sample_rows * (cuuint64_t)g.lda * 4};
        cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)t.box_l, (cuuint32_t)t.box_b};
        cuuint32_t estr[3] = {1, 1, 1};
        const float* basep = g.A + (down ? (int64_t)tap * g.lda : 0);
        CUresult r = enc(&tmAs[tap], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(basep), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MUGD_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled(A) failed with %d (K=%d L=%d B=%d lda=%lld)", (int)r, g.K,
                     t.Lrows, t.Bs, (long long)g.lda);
    }
    for (int w = 0; w < 2; ++w) {
        const cuuint64_t ktot = (cuuint64_t)g.taps * g.K;
        cuuint64_t dims[2] = {ktot, (cuuint64_t)g.N};
        cuuint64_t strides[1] = {ktot * 4};
        cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)(t.BN / t.mc)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(w == 0 ? &tmWhi : &tmWlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w == 0 ? g.W_hi : g.W_lo), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MUGD_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    TcParams p;
    p.g = g;
    p.ws = (float*)g.workspace;
    p.counters = g.counters;
    p.splits = t.splits;
    p.total_it = t.total_it;
    p.kblocks = g.K / TC_BK;
    p.Lrows = t.Lrows;
    p.Bs = t.Bs;
    p.box_l = t.box_l;
    p.box_b = t.box_b;
    p.tiles_per_sample = t.tiles_per_sample;
    p.dbg = g_tc_dbg;
    p.dbg_plain_store = g_tc_plain_store;
    p.pdl_reduce = (g_tc_pdl_reduce && t.splits > 1 && t.mc == 1 && !(g_tc_cluster && t.splits > 1)) ? 1 : 0;
    p.single_pass = g_tc_single_pass ? 1 : 0;
    p.cluster = use_cluster ? 1 : 0;
    p.inkernel_reduce = (!use_cluster && t.splits > 1 && t.splits <= g_tc_inkernel_max && g.counters && g.n_counters >= t.gx * t.gy) ? 1 : 0;
    if (g_tc_coop_reduce && !use_cluster && t.splits > 1 && t.mc == 1 && g.counters && g.n_counters >= t.gx * t.gy &&
        t.gx * t.gy * t.splits <= dev.sm_count)
        p.inkernel_reduce = 2;
    int rc;
#define TC_GO(BN_, AT_, MC_) rc = tc_launch<BN_, AT_, MC_>(tmAs, tmWhi, tmWlo, p, t, st)
    if (!g_tc_a_in_tmem) {
        if (t.BN == 256) TC_GO(256, false, 1); else if (t.BN == 128) TC_GO(128, false, 1); else TC_GO(64, false, 1);
    } else if (t.mc == 2) {
        if (t.BN == 256) TC_GO(256, true, 2); else TC_GO(128, true, 2);
    } else if (t.mc == 4) {
        if (t.BN == 256) TC_GO(256, true, 4); else TC_GO(128, true, 4);
    } else {
        if (t.BN == 256) TC_GO(256, true, 1); else if (t.BN == 128) TC_GO(128, true, 1); else TC_GO(64, true, 1);
    }
#undef TC_GO
    if (rc != MUGD_OK) return rc;
    if (launches) *launches += (t.splits > 1 && !use_cluster && !p.inkernel_reduce) ? 2 : 1;
    return MUGD_OK;
}

}  // namespace mugd

extern "C" int mugd_set_tc_a_in_tmem(int enabled) {
    mugd::g_tc_a_in_tmem = enabled != 0;
    return MUGD_OK;
}

extern "C" int mugd_set_tc_inkernel_reduce_max(int max_splits) {
    mugd::g_tc_inkernel_max = max_splits < 0 ? 0 : max_splits;
    return MUGD_OK;
}

extern "C" int mugd_set_tc_single_pass_tf32(int enabled) {
    mugd::g_tc_single_pass = enabled != 0;
    return MUGD_OK;
}

extern "C" int mugd_set_tc_narrow_tiles(int enabled, float kstep_us) {
    mugd::g_tc_narrow_tiles = enabled ? 1 : 0;
    if (kstep_us > 0.f) mugd::g_tc_kstep64 = kstep_us;
    return MUGD_OK;
}

extern "C" int mugd_debug_set_tc_cost(float kstep128_us, float kstep256_us, float split_us) {
    if (kstep128_us > 0.f) mugd::g_tc_cost[0] = kstep128_us;
    if (kstep256_us > 0.f) mugd::g_tc_cost[1] = kstep256_us;
    if (split_us > 0.f) mugd::g_tc_cost[2] = split_us;
    return MUGD_OK;
}

extern "C" int mugd_set_tc_coop_reduce(int enabled, float split_cost_us) {
    mugd::g_tc_coop_reduce = enabled ? 1 : 0;
    if (split_cost_us > 0.f) mugd::g_tc_coop_cost = split_cost_us;
    return MUGD_OK;
}

extern "C" int mugd_debug_set_tc_plain_store(int enabled) {
    mugd::g_tc_plain_store = enabled ? 1 : 0;
    return MUGD_OK;
}

extern "C" int mugd_set_tc_pdl_reduce(int enabled) {
    mugd::g_tc_pdl_reduce = enabled ? 1 : 0;
    return MUGD_OK;
}

extern "C" int mugd_set_tc_multicast(int max_cluster) {
    mugd::g_tc_multicast = (max_cluster == 2 || max_cluster == 4) ? max_cluster : 0;
    return MUGD_OK;
}

extern "C" int mugd_debug_set_tc_tile_n(int bn) {
    mugd::g_tc_force_bn = (bn == 64 || bn == 128 || bn == 256) ? bn : 0;
    return MUGD_OK;
}

extern "C" int mugd_set_tc_cluster_reduce(int enabled) {
    mugd::g_tc_cluster = enabled != 0;
    return MUGD_OK;
}

extern "C" int mugd_debug_set_tc_timing(long long* device_buf4) {
    mugd::g_tc_dbg = device_buf4;
    return MUGD_OK;
}

extern "C" int mugd_gemm_tc_query(mugd_handle*, const mugd_gemm* g, int32_t sm_count, int32_t* supported, int32_t* splits,
                                  int64_t* workspace_bytes, int32_t* n_tiles) {
    using namespace mugd;
    MUGD_REQUIRE(g, "gemm_tc_query: null");
    const bool ok = (g->conv_mode == MUGD_CONV_NONE || g->conv_mode == MUGD_CONV_SAME || g->conv_mode == MUGD_CONV_DOWN ||
                     g->conv_mode == MUGD_CONV_TAPS) && g->K % TC_BK == 0 && g->N >= 64 && g->N % 4 == 0;
    if (supported) *supported = ok ? 1 : 0;
    if (!ok) {
        if (splits) *splits = 0;
        if (workspace_bytes) *workspace_bytes = 0;
        if (n_tiles) *n_tiles = 0;
        return MUGD_OK;
    }
    const TcGeometry t = tc_geometry(*g, sm_count > 0 ? sm_count : 148, g->split_k);
    if (splits) *splits = t.splits;
    if (workspace_bytes) *workspace_bytes = t.ws_floats * 4;
    if (n_tiles) *n_tiles = t.gx * t.gy;
    return MUGD_OK;
}
