// Attention of mug/model/attention.py:91-126 (CrossAttention.forward) with both contractions on the tcgen05 tensor
// cores, fp32 in / fp32 out through the same 3xTF32 split as gemm_tc.cu:
//
//     idx_ij = clamp(j - i, -P, P) + P
//     s_ij   = (q_i . k_j + relpos[idx_ij, h]) * scale                 S = Q K^T   : tcgen05.mma, A = Q from TMEM
//     o_i    = sum_j softmax_j(s_i)_j * cgain[idx_ij, h] * v_j         O = P V     : tcgen05.mma, A = P from TMEM
//
// One CTA owns 128 queries of one (sample, head) and streams 128-key tiles (flash style: no [Lq, Lk] matrix in
// memory, running max / sum per query row).  Thread t and thread t+128 share query row (t & 127) = TMEM lane:
//   * Q is split into q_hi / q_lo once and lives in tensor memory for the whole CTA (A operand, "TS" form);
//   * per key tile, threads 0-127 split K rows into k_hi / k_lo (K-major, SWIZZLE_128B canonical layout written by
//     hand, B operand of S = Q K^T) while threads 128-255 write V transposed (V^T hi / lo: rows = head channels,
//     keys contiguous = K-major B operand of O = P V);
//   * S lands in TMEM columns [0,128); each thread reads its half of the row 16 columns at a time, applies the
//     relative-position bias, scale and key mask, and the two halves combine max / sum through shared memory;
//   * P * gain is split into hi / lo and written back to tensor memory (hi over the S columns it came from, lo next
//     to it): the second MMA consumes it from there, so P never touches shared or global memory;
//   * the per-tile O lands in TMEM and is folded into the register accumulator with the usual exp(m_old - m_new).
// The FFMA kernel in attention.cu stays as the exact-fp32 referee (mugd_set_attention_impl(0)).
#include "common.cuh"

#include <math.h>

namespace mugd {
namespace atc {

constexpr int THREADS = 256;
constexpr int BQ = 128;
constexpr int BKV = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (same encoding as gemm_tc.cu): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// 8 consecutive 32-bit TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

template <int D>
struct Smem {
    static constexpr int KSLABS = (D + 31) / 32;
    static constexpr uint32_t K_SLAB = BKV * 128;          // 128 keys x (32 fp32 = 128 B)
    static constexpr uint32_t K_BYTES = KSLABS * K_SLAB;   // one of k_hi / k_lo
    static constexpr uint32_t V_SLAB = D * 128;            // D channel rows x (32 keys = 128 B)
    static constexpr uint32_t V_BYTES = (BKV / 32) * V_SLAB;
    static constexpr uint32_t TILE_BYTES = 2 * K_BYTES + 2 * V_BYTES;
    static constexpr uint32_t AUX_BYTES = 64 + 4 * BQ * 4;  // 2 mbarriers + tmem slot | mx[2][128] rs[2][128]
    static size_t total(int pos_max) { return TILE_BYTES + AUX_BYTES + 2 * (2 * pos_max + 1) * 4 + 1024; }
};

template <int D>
__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const mugd_attention a) {
    using S = Smem<D>;
    constexpr int HC = D / 2;                               // Q / O columns owned by one thread of a row pair
    constexpr uint32_t TM_S = 0, TM_PLO = 128, TM_O = 256, TM_QHI = 320, TM_QLO = 384, TM_COLS = 512;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;           // SWIZZLE_128B operands need 1024-byte alignment
    const uint32_t k_hi = base, k_lo = k_hi + S::K_BYTES, v_hi = k_lo + S::K_BYTES, v_lo = v_hi + S::V_BYTES;
    const uint32_t aux = v_lo + S::V_BYTES;
    const uint32_t bar_s = aux, bar_o = aux + 8, tmem_slot = aux + 16;
    float* red = reinterpret_cast<float*>(smem_raw + (aux - raw) + 64);   // mx[2][128], rs[2][128]
    float* rel = red + 4 * BQ;
    const int P = a.pos_max, NT = 2 * P + 1;
    float* cg = rel + NT;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int g = tid >> 7, r = tid & 127;                  // thread group (column half), query row = TMEM lane
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;

    pdl_trigger();
    if (tid == 0) {
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    pdl_wait();

    for (int t = tid; t < NT; t += THREADS) {
        rel[t] = a.relpos[t * a.H + h];
        cg[t] = a.cgain[t * a.H + h];
    }
    // ---- Q row half -> q_hi / q_lo in tensor memory --------------------------------------------------------
    const int qi = q0 + r;
    {
        const float* qp = a.q + ((int64_t)b * a.Lq + qi) * a.ldq + h * D + g * HC;
#pragma unroll
        for (int c = 0; c < HC / 8; ++c) {
            float hi[8], lo[8];
            float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
            if (qi < a.Lq) { x0 = ld_f4(qp + c * 8); x1 = ld_f4(qp + c * 8 + 4); }
            const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) { hi[j] = to_tf32(x[j]); lo[j] = to_tf32(x[j] - hi[j]); }
            tmem_st8(lane_addr + TM_QHI + g * HC + c * 8, hi);
            tmem_st8(lane_addr + TM_QLO + g * HC + c * 8, lo);
        }
        tmem_wait_st();
    }
    float m_i = -INFINITY, l_i = 0.f, o[HC];
#pragma unroll
    for (int c = 0; c < HC; ++c) o[c] = 0.f;
    // instruction descriptor: D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2, K-major A/B, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BQ >> 4) << 24);
    uint32_t ph = 0;

    for (int j0 = 0; j0 < a.Lk; j0 += BKV) {
        const int nk = min(BKV, a.Lk - j0);
        const int NK = (nk + 15) & ~15;                     // MMA N (S) and K extent (PV): padded keys are zero / masked
        // ---- stage this key tile: every MMA that read the previous one has retired (bar_o wait below) -------
        if (r < NK) {
            if (g == 0) {
                const float* kp = a.k + ((int64_t)b * a.Lk + j0 + r) * a.ldk + h * D;
#pragma unroll
                for (int c = 0; c < D / 4; ++c) {
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nk) x = ld_f4(kp + c * 4);
                    float4 hi, lo;
                    hi.x = to_tf32(x.x); hi.y = to_tf32(x.y); hi.z = to_tf32(x.z); hi.w = to_tf32(x.w);
                    lo.x = to_tf32(x.x - hi.x); lo.y = to_tf32(x.y - hi.y); lo.z = to_tf32(x.z - hi.z); lo.w = to_tf32(x.w - hi.w);
                    const uint32_t off = (uint32_t)(c >> 3) * S::K_SLAB + (uint32_t)r * 128u + (uint32_t)(((c & 7) ^ (r & 7)) << 4);
                    sts_f4(k_hi + off, hi);
                    sts_f4(k_lo + off, lo);
                }
            } else {
                const float* vp = a.v + ((int64_t)b * a.Lk + j0 + r) * a.ldv + h * D;
                const uint32_t col = (uint32_t)(r >> 5) * S::V_SLAB + (uint32_t)((r & 3) << 2);
                const int kc = (r & 31) >> 2;               // 16-byte chunk of this key inside its 32-key slab row
#pragma unroll
                for (int c = 0; c < D / 4; ++c) {
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nk) x = ld_f4(vp + c * 4);
                    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int d = c * 4 + j;
                        const float hi = to_tf32(xs[j]), lo = to_tf32(xs[j] - hi);
                        const uint32_t off = col + (uint32_t)d * 128u + (uint32_t)((kc ^ (d & 7)) << 4);
                        sts_f1(v_hi + off, hi);
                        sts_f1(v_lo + off, lo);
                    }
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA
        fence_before();
        __syncthreads();
        // ---- S = Q K^T ----------------------------------------------------------------------------------------
        if (tid == 0) {
            fence_after();
            const uint32_t idesc = idesc0 | ((uint32_t)(NK >> 3) << 17);
#pragma unroll
            for (int kk = 0; kk < D / 8; ++kk) {
                const uint32_t so = (uint32_t)(kk >> 2) * S::K_SLAB;
                const uint64_t ko = (uint64_t)((kk & 3) * 2);          // 8 fp32 = 32 bytes = 2 x 16-byte units
                const uint64_t dh = umma_desc(k_hi + so) + ko, dl = umma_desc(k_lo + so) + ko;
                umma_tf32_ts(tmem_base + TM_S, tmem_base + TM_QLO + kk * 8, dh, idesc, kk > 0 ? 1u : 0u);
                umma_tf32_ts(tmem_base + TM_S, tmem_base + TM_QHI + kk * 8, dl, idesc, 1u);
                umma_tf32_ts(tmem_base + TM_S, tmem_base + TM_QHI + kk * 8, dh, idesc, 1u);
            }
            umma_commit(bar_s);
        }
        mbar_wait(bar_s, ph);
        fence_after();
        // ---- bias, scale, mask, online softmax on this thread's half of the row ---------------------------------
        const int split = min(NK, ((NK >> 4) + 1) / 2 * 16);
        const int c_lo = g == 0 ? 0 : split, c_hi = g == 0 ? split : NK;
        float mx = -INFINITY;
        for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            float v[16];
            tmem_ld8(lane_addr + TM_S + c0, v);
            tmem_ld8(lane_addr + TM_S + c0 + 8, v + 8);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int kj = j0 + c0 + j;
                const int idx = max(-P, min(P, kj - qi)) + P;
                const float s = (kj < a.Lk) ? (v[j] + rel[idx]) * a.scale : -INFINITY;
                mx = fmaxf(mx, s);
            }
        }
        red[g * BQ + r] = mx;
        __syncthreads();
        const float mnew = fmaxf(m_i, fmaxf(red[r], red[BQ + r]));     // finite: key j0 is always valid
        const float corr = expf(m_i - mnew);
        float rs = 0.f;
        for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            float v[16], phi[16], plo[16];
            tmem_ld8(lane_addr + TM_S + c0, v);
            tmem_ld8(lane_addr + TM_S + c0 + 8, v + 8);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int kj = j0 + c0 + j;
                const int idx = max(-P, min(P, kj - qi)) + P;
                const float s = (kj < a.Lk) ? (v[j] + rel[idx]) * a.scale : -INFINITY;
                const float pe = expf(s - mnew);                        // 0 for masked keys
                rs += pe;
                const float pg = pe * cg[idx];
                phi[j] = to_tf32(pg);
                plo[j] = to_tf32(pg - phi[j]);
            }
            tmem_st8(lane_addr + TM_S + c0, phi);                       // p_hi overwrites the logits it was made from
            tmem_st8(lane_addr + TM_S + c0 + 8, phi + 8);
            tmem_st8(lane_addr + TM_PLO + c0, plo);
            tmem_st8(lane_addr + TM_PLO + c0 + 8, plo + 8);
        }
        tmem_wait_st();
        red[2 * BQ + g * BQ + r] = rs;
        fence_before();
        __syncthreads();
        l_i = l_i * corr + (red[2 * BQ + r] + red[3 * BQ + r]);
        m_i = mnew;
        // ---- O_tile = P V ----------------------------------------------------------------------------------------
        if (tid == 0) {
            fence_after();
            const uint32_t idesc = idesc0 | ((uint32_t)(D >> 3) << 17);
            for (int kk = 0; kk < NK / 8; ++kk) {
                const uint32_t so = (uint32_t)(kk >> 2) * S::V_SLAB;
                const uint64_t ko = (uint64_t)((kk & 3) * 2);
                const uint64_t dh = umma_desc(v_hi + so) + ko, dl = umma_desc(v_lo + so) + ko;
                umma_tf32_ts(tmem_base + TM_O, tmem_base + TM_PLO + kk * 8, dh, idesc, kk > 0 ? 1u : 0u);
                umma_tf32_ts(tmem_base + TM_O, tmem_base + TM_S + kk * 8, dl, idesc, 1u);
                umma_tf32_ts(tmem_base + TM_O, tmem_base + TM_S + kk * 8, dh, idesc, 1u);
            }
            umma_commit(bar_o);
        }
        mbar_wait(bar_o, ph);
        fence_after();
#pragma unroll
        for (int c = 0; c < HC / 8; ++c) {
            float ot[8];
            tmem_ld8(lane_addr + TM_O + g * HC + c * 8, ot);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) o[c * 8 + j] = fmaf(o[c * 8 + j], corr, ot[j]);
        }
        ph ^= 1u;
        fence_before();                                     // orders these TMEM reads before the next tile's MMAs
    }
    if (qi < a.Lq) {
        const float inv = 1.0f / l_i;
        float* op = a.o + ((int64_t)b * a.Lq + qi) * a.ldo + h * D + g * HC;
#pragma unroll
        for (int c = 0; c < HC / 4; ++c)
            st_f4(op + c * 4, make_float4(o[c * 4] * inv, o[c * 4 + 1] * inv, o[c * 4 + 2] * inv, o[c * 4 + 3] * inv));
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TM_COLS) : "memory");
    }
}

template <int D>
static int launch(const mugd_attention& a, cudaStream_t st) {
    const size_t bytes = Smem<D>::total(a.pos_max);
    static size_t configured = 0;
    if (bytes > configured) {
        MUGD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = bytes;
    }
    dim3 grid((a.Lq + BQ - 1) / BQ, a.H, a.B);
    MUGD_CHECK_CUDA(launch_k(attention_tc_kernel<D>, grid, dim3(THREADS), bytes, st, a));
    return MUGD_OK;
}

}  // namespace atc

int launch_attention_tc(const DeviceInfo&, const mugd_attention& a, cudaStream_t st) {
    return (a.D == 32) ? atc::launch<32>(a, st) : (a.D == 48) ? atc::launch<48>(a, st) : atc::launch<64>(a, st);
}

}  // namespace mugd
