// Attention of mug/model/attention.py:91-126 (CrossAttention.forward), fp32, flash-style (no [Lq,Lk]
// matrix in memory):
//     idx_ij = clamp(j - i, -P, P) + P
//     s_ij   = (q_i . k_j + relpos[idx_ij, h]) * scale
//     o_i    = sum_j softmax_j(s_i)_j * cgain[idx_ij, h] * v_j
// The post-softmax gain multiplies the numerator only; the softmax denominator accumulates plain p.
// Self attention (Lk = Lq) and cross attention to the 21 prompt tokens (Lk = 21) share the kernel.
//
// One thread owns one query row (q, accumulator and running max/sum in registers); a CTA of 64 queries
// streams K/V tiles of 32 keys through shared memory, read back as warp-broadcast float4 (conflict-free).
#include "common.cuh"

#include <math.h>

namespace mugd {

constexpr int AT_Q = 64;    // queries (threads) per CTA
constexpr int AT_TK = 32;   // keys per smem tile

template <int D>
__global__ void __launch_bounds__(AT_Q)
attention_kernel(const mugd_attention a) {
    __shared__ __align__(16) float Ks[AT_TK][D];
    __shared__ __align__(16) float Vs[AT_TK][D];
    extern __shared__ float tabs[];           // [2][2P+1] : relpos column h, cgain column h
    const int P = a.pos_max, NT = 2 * P + 1;
    float* rel = tabs;
    float* cg = tabs + NT;

    const int b = blockIdx.z, h = blockIdx.y;
    const int i = blockIdx.x * AT_Q + threadIdx.x;    // query index
    const bool active = i < a.Lq;
    for (int t = threadIdx.x; t < NT; t += AT_Q) {
        rel[t] = a.relpos[t * a.H + h];
        cg[t] = a.cgain[t * a.H + h];
    }

    float q[D], acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    if (active) {
        const float* qp = a.q + ((int64_t)b * a.Lq + i) * a.ldq + h * D;
#pragma unroll
        for (int d = 0; d < D; d += 4) {
            const float4 v = ld_f4(qp + d);
            q[d] = v.x; q[d + 1] = v.y; q[d + 2] = v.z; q[d + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int d = 0; d < D; ++d) q[d] = 0.f;
    }
    float mrun = -INFINITY, lrun = 0.f;

    const float* kbase = a.k + (int64_t)b * a.Lk * a.ldk + h * D;
    const float* vbase = a.v + (int64_t)b * a.Lk * a.ldv + h * D;
    constexpr int QD = D / 4;
    for (int j0 = 0; j0 < a.Lk; j0 += AT_TK) {
        __syncthreads();   // previous tile fully consumed (also orders the table writes on the first pass)
        for (int t = threadIdx.x; t < AT_TK * QD; t += AT_Q) {
            const int r = t / QD, c = (t - r * QD) * 4;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (j0 + r < a.Lk) {
                kv = ld_f4(kbase + (int64_t)(j0 + r) * a.ldk + c);
                vv = ld_f4(vbase + (int64_t)(j0 + r) * a.ldv + c);
            }
            *reinterpret_cast<float4*>(&Ks[r][c]) = kv;
            *reinterpret_cast<float4*>(&Vs[r][c]) = vv;
        }
        __syncthreads();
        const int nk = min(AT_TK, a.Lk - j0);
        float s[AT_TK];
        float tmax = -INFINITY;
#pragma unroll
        for (int r = 0; r < AT_TK; ++r) {
            float dot = 0.f;
#pragma unroll
            for (int d = 0; d < D; d += 4) {
                const float4 kv = *reinterpret_cast<const float4*>(&Ks[r][d]);
                dot = fmaf(q[d], kv.x, dot);
                dot = fmaf(q[d + 1], kv.y, dot);
                dot = fmaf(q[d + 2], kv.z, dot);
                dot = fmaf(q[d + 3], kv.w, dot);
            }
            int idx = (j0 + r) - i;
            idx = max(-P, min(P, idx)) + P;
            const float sv = (r < nk) ? (dot + rel[idx]) * a.scale : -INFINITY;
            s[r] = sv;
            tmax = fmaxf(tmax, sv);
        }
        const float mnew = fmaxf(mrun, tmax);          // finite: every tile holds >= 1 valid key
        const float corr = expf(mrun - mnew);          // exp(-inf) = 0 on the first tile
        lrun *= corr;
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] *= corr;
#pragma unroll
        for (int r = 0; r < AT_TK; ++r) {
            const float pexp = expf(s[r] - mnew);      // 0 for masked keys
            lrun += pexp;
            int idx = (j0 + r) - i;
            idx = max(-P, min(P, idx)) + P;
            const float pc = pexp * cg[idx];
#pragma unroll
            for (int d = 0; d < D; d += 4) {
                const float4 vv = *reinterpret_cast<const float4*>(&Vs[r][d]);
                acc[d] = fmaf(pc, vv.x, acc[d]);
                acc[d + 1] = fmaf(pc, vv.y, acc[d + 1]);
                acc[d + 2] = fmaf(pc, vv.z, acc[d + 2]);
                acc[d + 3] = fmaf(pc, vv.w, acc[d + 3]);
            }
        }
        mrun = mnew;
    }
    if (active) {
        const float inv = 1.0f / lrun;
        float* op = a.o + ((int64_t)b * a.Lq + i) * a.ldo + h * D;
#pragma unroll
        for (int d = 0; d < D; d += 4)
            st_f4(op + d, make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv));
    }
}

int launch_attention(const DeviceInfo&, const mugd_attention& a, cudaStream_t st, int* launches) {
    MUGD_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0, "attention: empty shape");
    MUGD_REQUIRE(a.D == 32 || a.D == 48 || a.D == 64, "attention: head dim %d not in {32,48,64}", a.D);
    MUGD_REQUIRE(a.pos_max >= 0 && a.pos_max <= 1024, "attention: pos_max %d", a.pos_max);
    MUGD_REQUIRE(aligned16(a.q) && aligned16(a.k) && aligned16(a.v) && aligned16(a.o) && a.ldq % 4 == 0 && a.ldk % 4 == 0 &&
                     a.ldv % 4 == 0 && a.ldo % 4 == 0, "attention: alignment");
    MUGD_REQUIRE(a.ldq >= a.H * a.D && a.ldk >= a.H * a.D && a.ldv >= a.H * a.D && a.ldo >= a.H * a.D, "attention: ld < H*D");
    MUGD_REQUIRE(a.relpos && a.cgain, "attention: tables missing");
    dim3 grid((a.Lq + AT_Q - 1) / AT_Q, a.H, a.B);
    const size_t dyn = sizeof(float) * 2 * (2 * a.pos_max + 1);
    if (a.D == 32) attention_kernel<32><<<grid, AT_Q, dyn, st>>>(a);
    else if (a.D == 48) attention_kernel<48><<<grid, AT_Q, dyn, st>>>(a);
    else attention_kernel<64><<<grid, AT_Q, dyn, st>>>(a);
    MUGD_CHECK_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    return MUGD_OK;
}

}  // namespace mugd
