// S4 layer pieces (mug/model/s4.py):
//
// (1) s4conv: the per-step part of S4.forward (s4.py:1503-1532).  The reference multiplies rfft(u, 2L)
//     by rfft(K, 2L) and keeps the first L samples of the inverse -- i.e. the causal convolution
//         y[b, l, h] = sum_{j <= l} K[h, j] * u[b, l - j, h]
//     which is evaluated here directly in fp32 (agrees with the FFT form to ~4e-7), fused
//     with the D*u skip (s4.py:1514) and the exact-erf GELU (s4.py:1532).
//
// (2) s4 kernel generation: SSKernelNPLR.forward (s4.py:706-832) for rank 1 / channels 1 / rate 1 /
//     no state, with the NON-conjugate Cauchy sum `cauchy_naive` (s4.py:140-147) the reference falls back
//     to.  The reference regenerates K on every forward although it depends on parameters only; here it
//     runs once per (model, L) in fp64 and is checked tap-for-tap against the reference's K.
#include "common.cuh"

#include <math.h>

namespace mugd {

// =====================================================================================================
// (1) causal long convolution.  lane = channel (coalesced), each warp owns blocks of 8 consecutive
// outputs; a 16-deep register window slides over u (smem) while K taps stream from L1/L2.
// =====================================================================================================
constexpr int S4_WARPS = 8;
constexpr int S4_R = 8;         // outputs per block

// CH channels per CTA (32: one lane per channel; 16: the two half-warps share 16 channels and split the taps of every
// chunk by parity, accumulators are added with one shuffle at the end -- used when two L x 32 tiles do not fit).
// KS: kernel taps staged in shared memory next to u (2 * L * CH * 4 B); otherwise streamed through L1 (very long L).
template <bool KS, int CH>
__global__ void __launch_bounds__(32 * S4_WARPS)
s4conv_kernel(const mugd_s4conv s, int nsplit) {
    constexpr int NSUB = 32 / CH;
    extern __shared__ float smem_s4[];     // us[L][CH] then ks[L][CH]
    float* us = smem_s4;
    float* ks = smem_s4 + (size_t)s.L * CH;
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = lane % CH, sub = lane / CH;
    const int h = blockIdx.x * CH + ch;
    const int b = blockIdx.y;
    const int L = s.L;
    const float* ub = s.u + (int64_t)b * L * s.ldu + blockIdx.x * CH;
    const float* Kc = s.Kt + blockIdx.x * CH;  // tap j of channel c at Kc[j*H + c]
    for (int i = threadIdx.x; i < L * CH; i += 32 * S4_WARPS) {
        const int l = i / CH, c = i - l * CH;
        us[i] = ub[(int64_t)l * s.ldu + c];
        if (KS) ks[i] = Kc[(int64_t)l * s.H + c];
    }
    __syncthreads();

    const float* Kh = s.Kt + h;
    const float Dh = s.D[h];
    float* yb = s.y + (int64_t)b * L * s.ldy + h;
    const int nblk = (L + S4_R - 1) / S4_R;
    const int npairs = (nblk + 1) / 2;
    const int worker = blockIdx.z * S4_WARPS + warp;
    const int nworkers = nsplit * S4_WARPS;
    // the cost of block bi grows linearly with bi (causal): pairing block p with block nblk-1-p gives every
    // worker the same amount of work
    for (int p = worker; p < npairs; p += nworkers) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int bi = half == 0 ? (nblk - 1 - p) : p;
            if (half == 1 && bi == nblk - 1 - p) break;
            const int l0 = bi * S4_R;
            float acc[S4_R];
#pragma unroll
            for (int r = 0; r < S4_R; ++r) acc[r] = 0.f;
            float win[2 * S4_R];               // win[i] = u[l0 - jc - 8 + i]
#pragma unroll
            for (int r = 0; r < S4_R; ++r) {
                const int li = l0 + r;
                win[S4_R + r] = (li < L) ? us[li * CH + ch] : 0.f;
            }
            for (int jc = 0; jc < l0 + S4_R; jc += S4_R) {
                float kk[S4_R / NSUB];
#pragma unroll
                for (int r = 0; r < S4_R; ++r) {
                    const int li = l0 - jc - S4_R + r;
                    win[r] = (li >= 0) ? us[li * CH + ch] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < S4_R / NSUB; ++i) {
                    const int j = jc + sub + NSUB * i;
                    kk[i] = (j < L) ? (KS ? ks[j * CH + ch] : __ldg(Kh + (int64_t)j * s.H)) : 0.f;
                }
#pragma unroll
                for (int i = 0; i < S4_R / NSUB; ++i) {
#pragma unroll
                    for (int r = 0; r < S4_R; ++r) {
                        // tap jj = sub + NSUB*i : window index S4_R + r - jj.  `sub` is 0 when NSUB == 1, else 0/1: select
                        const float wv = (NSUB == 1) ? win[S4_R + r - i] : (sub ? win[S4_R + r - 1 - NSUB * i] : win[S4_R + r - NSUB * i]);
                        acc[r] = fmaf(kk[i], wv, acc[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < S4_R; ++r) win[S4_R + r] = win[r];
            }
            if (NSUB == 2) {
#pragma unroll
                for (int r = 0; r < S4_R; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 16);
            }
            if (sub == 0) {
#pragma unroll
                for (int r = 0; r < S4_R; ++r) {
                    const int li = l0 + r;
                    if (li < L) yb[(int64_t)li * s.ldy] = gelu_f(acc[r] + Dh * us[li * CH + ch]);
                }
            }
        }
    }
}

int launch_s4conv(const DeviceInfo& dev, const mugd_s4conv& s, cudaStream_t st, int* launches) {
    MUGD_REQUIRE(s.B > 0 && s.L > 0 && s.H > 0 && s.H % 32 == 0, "s4conv: H=%d must be a positive multiple of 32", s.H);
    MUGD_REQUIRE(s.ldu >= s.H && s.ldy >= s.H, "s4conv: ld < H");
    const size_t tile32 = (size_t)s.L * 32 * sizeof(float);
    MUGD_REQUIRE((int)tile32 <= dev.max_smem_optin, "s4conv: L=%d needs %zu B of shared memory (max %d)", s.L, tile32, dev.max_smem_optin);
    // 0: u+K tiles of 32 channels; 1: u+K tiles of 16 channels; 2: u tile of 32 channels, K through L1
    const int variant = (2 * tile32 <= (size_t)dev.max_smem_optin) ? 0 : (tile32 <= (size_t)dev.max_smem_optin ? 1 : 2);
    static bool configured = false;
    if (!configured) {
        MUGD_CHECK_CUDA(cudaFuncSetAttribute(s4conv_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.max_smem_optin));
        MUGD_CHECK_CUDA(cudaFuncSetAttribute(s4conv_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.max_smem_optin));
        MUGD_CHECK_CUDA(cudaFuncSetAttribute(s4conv_kernel<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.max_smem_optin));
        configured = true;
    }
    const int ch = variant == 1 ? 16 : 32;
    const int base = (s.H / ch) * s.B;
    const int npairs = ((s.L + S4_R - 1) / S4_R + 1) / 2;
    int nsplit = 1;
    while (base * nsplit < 2 * dev.sm_count && nsplit * 2 * S4_WARPS <= npairs && nsplit < 16) nsplit *= 2;
    dim3 grid(s.H / ch, s.B, nsplit);
    if (variant == 0) MUGD_CHECK_CUDA(launch_k(s4conv_kernel<true, 32>, grid, dim3(32 * S4_WARPS), 2 * tile32, st, s, nsplit));
    else if (variant == 1) MUGD_CHECK_CUDA(launch_k(s4conv_kernel<true, 16>, grid, dim3(32 * S4_WARPS), tile32, st, s, nsplit));
    else MUGD_CHECK_CUDA(launch_k(s4conv_kernel<false, 32>, grid, dim3(32 * S4_WARPS), tile32, st, s, nsplit));
    if (launches) *launches += 1;
    return MUGD_OK;
}

// =====================================================================================================
// (2) kernel generation (fp64).  With w' = w*dt, omega_f = exp(-2 pi i f / L):
//   reference:  z = 2(1-omega)/(1+omega);  r_xy = dt * sum_n v_xy[n] / (z - w'_n)
//               k_f = (r00 - r01 r10 / (1 + r11)) * 2 / (1 + omega);   K = irfft(k_f, L)[:L_out]
//   Multiplying numerator and denominator by (1+omega) -- an identity for ANY omega, exact or not --
//   removes the 0/0 at the Nyquist node (where the reference relies on rounding noise of its complex64
//   omega^f):
//               s_xy = dt * sum_n v_xy[n] / (2(1-omega) - w'_n (1+omega)),     r_xy = (1+omega) s_xy
//               k_f  = 2 * ( s00 - (1+omega) s01 s10 / (1 + (1+omega) s11) )
//   v00 = B*C, v01 = B*conj(P), v10 = P*C, v11 = P*conj(P)              (s4.py:771-778)
// =====================================================================================================
struct cd { double re, im; };
__device__ __forceinline__ cd cmul(cd a, cd b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cd cadd(cd a, cd b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cd csub(cd a, cd b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cd cdiv(cd a, cd b) {
    const double d = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

__global__ void s4_kf_kernel(const float* __restrict__ log_dt, const float* __restrict__ Bri,
                             const float* __restrict__ Cri, const float* __restrict__ Pri,
                             const float* __restrict__ inv_w_real, const float* __restrict__ w_imag,
                             const float* __restrict__ omega_ri, int H, int N, int Lint, double2* __restrict__ kf) {
    const int nf = Lint / 2 + 1;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = blockIdx.y;
    if (f >= nf) return;
    const double dt = exp((double)log_dt[h]);
    // FFT nodes: the caller's table (the reference evaluates omega^f as a complex64 power, s4.py:595-598,
    // which is off by up to ~5e-6 at f = L/2; using the same nodes reproduces its kernel to ~2e-6) or exact.
    cd om;
    if (omega_ri) {
        om = {(double)omega_ri[2 * f], (double)omega_ri[2 * f + 1]};
    } else {
        double sn, cs;
        sincospi(-2.0 * (double)f / (double)Lint, &sn, &cs);
        om = {cs, sn};
    }
    const cd one_m = {2.0 * (1.0 - om.re), -2.0 * om.im};     // 2(1-omega)
    const cd one_p = {1.0 + om.re, om.im};                    // 1+omega
    cd s00 = {0, 0}, s01 = {0, 0}, s10 = {0, 0}, s11 = {0, 0};
    for (int n = 0; n < N; ++n) {
        const int64_t o = ((int64_t)h * N + n);
        const cd w = {-exp((double)inv_w_real[o]) * dt, (double)w_imag[o] * dt};
        const cd Bc = {(double)Bri[2 * o], (double)Bri[2 * o + 1]};
        const cd Cc = {(double)Cri[2 * o], (double)Cri[2 * o + 1]};
        const cd Pc = {(double)Pri[2 * o], (double)Pri[2 * o + 1]};
        const cd Qc = {Pc.re, -Pc.im};
        const cd den = csub(one_m, cmul(w, one_p));
        const cd inv = cdiv({1.0, 0.0}, den);
        s00 = cadd(s00, cmul(cmul(Bc, Cc), inv));
        s01 = cadd(s01, cmul(cmul(Bc, Qc), inv));
        s10 = cadd(s10, cmul(cmul(Pc, Cc), inv));
        s11 = cadd(s11, cmul(cmul(Pc, Qc), inv));
    }
    s00 = {s00.re * dt, s00.im * dt}; s01 = {s01.re * dt, s01.im * dt};
    s10 = {s10.re * dt, s10.im * dt}; s11 = {s11.re * dt, s11.im * dt};
    const cd num = cmul(one_p, cmul(s01, s10));
    const cd den = cadd({1.0, 0.0}, cmul(one_p, s11));
    const cd k = csub(s00, cdiv(num, den));
    kf[(int64_t)h * nf + f] = make_double2(2.0 * k.re, 2.0 * k.im);
}

// inverse real DFT of length Lint (C2R semantics of torch.fft.irfft: imaginary parts of the DC and
// Nyquist bins are ignored), truncated to L_out taps, written tap-major Kt[l][h].
__global__ void s4_irfft_kernel(const double2* __restrict__ kf, int H, int Lint, int Lout, float* __restrict__ Kt) {
    extern __shared__ double2 sm[];       // [nf] spectrum of this h, then [Lint] twiddles
    const int nf = Lint / 2 + 1;
    double2* X = sm;
    double2* tw = sm + nf;
    const int h = blockIdx.x;
    for (int f = threadIdx.x; f < nf; f += blockDim.x) X[f] = kf[(int64_t)h * nf + f];
    for (int m = threadIdx.x; m < Lint; m += blockDim.x) {
        double sn, cs;
        sincospi(2.0 * (double)m / (double)Lint, &sn, &cs);
        tw[m] = make_double2(cs, sn);
    }
    __syncthreads();
    const bool even = (Lint % 2) == 0;
    const int fmax = even ? nf - 1 : nf;  // exclusive upper bound of the doubled interior bins
    for (int l = threadIdx.x; l < Lout; l += blockDim.x) {
        double acc = X[0].x;
        int ph = 0;
        for (int f = 1; f < fmax; ++f) {
            ph += l;
            if (ph >= Lint) ph -= Lint;
            acc += 2.0 * (X[f].x * tw[ph].x - X[f].y * tw[ph].y);
        }
        if (even) acc += (l & 1) ? -X[nf - 1].x : X[nf - 1].x;
        Kt[(int64_t)l * H + h] = (float)(acc / (double)Lint);
    }
}

}  // namespace mugd

extern "C" int mugd_s4_kernel_gen(mugd_handle*, const float* log_dt, const float* Bri, const float* Cri,
                                  const float* Pri, const float* inv_w_real, const float* w_imag,
                                  const float* omega_ri, int32_t H, int32_t N, int32_t L_internal, int32_t L_out,
                                  float* Kt, void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace mugd;
    MUGD_REQUIRE(H > 0 && N > 0 && L_internal > 0 && L_out > 0 && L_out <= L_internal,
                 "s4_kernel_gen: bad shape H=%d N=%d L_internal=%d L_out=%d (L_out must be <= L_internal; lengthen C~ with "
                 "the host-side setup first, s4.py:557-584)", H, N, L_internal, L_out);
    const int nf = L_internal / 2 + 1;
    MUGD_REQUIRE(workspace && workspace_bytes >= (int64_t)sizeof(double2) * H * nf, "s4_kernel_gen: workspace too small");
    MUGD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "s4_kernel_gen: workspace alignment");
    cudaStream_t st = (cudaStream_t)stream;
    double2* kf = (double2*)workspace;
    dim3 g1((nf + 127) / 128, H);
    s4_kf_kernel<<<g1, 128, 0, st>>>(log_dt, Bri, Cri, Pri, inv_w_real, w_imag, omega_ri, H, N, L_internal, kf);
    MUGD_CHECK_CUDA(cudaGetLastError());
    const size_t smem = sizeof(double2) * (size_t)(nf + L_internal);
    MUGD_REQUIRE(smem <= 200 * 1024, "s4_kernel_gen: L_internal=%d too long for the one-shot DFT", L_internal);
    if (smem > 48 * 1024)
        MUGD_CHECK_CUDA(cudaFuncSetAttribute(s4_irfft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    s4_irfft_kernel<<<H, 256, smem, st>>>(kf, H, L_internal, L_out, Kt);
    MUGD_CHECK_CUDA(cudaGetLastError());
    return MUGD_OK;
}
