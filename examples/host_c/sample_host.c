/* A host without Python: run one Mug-Diffusion sampling request (S DDIM steps with classifier-free guidance + first-stage decode)
 * from a bundle written by `python -m mug_diffusion_b200.bundle`, through the C ABI of libmugd.so only.
 *
 *   make -C examples/host_c            (gcc + the CUDA runtime; no torch, no Python)
 *   examples/host_c/sample_host <bundle dir>
 *
 * manifest.txt lines:  region <name> <bytes> zero|file <file>   |  plan <file> run|graph  |  sample <eval> <tail> <steps>
 *                      expect <region> <bytes> <file>           (outputs to compare; exit status 1 on mismatch)
 * This mirrors what DDIMSampler.sample + model.decode do in the reference (mug/diffusion/ddim.py:56-196, diffusion.py:49-50). */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mugd.h"

#define MAXR 128
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)
#define MK(x) do { int r_ = (x); if (r_ != MUGD_OK) { fprintf(stderr, "%s:%d libmugd status %d: %s\n", __FILE__, __LINE__, r_, mugd_last_error()); exit(2); } } while (0)

static mugd_region regions[MAXR];
static char names[MAXR][48];
static int n_regions = 0;

static void* read_file(const char* dir, const char* name, long long expect_bytes) {
    char path[1024];
    snprintf(path, sizeof(path), "%s/%s", dir, name);
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    void* buf = malloc((size_t)expect_bytes);
    if (fread(buf, 1, (size_t)expect_bytes, f) != (size_t)expect_bytes) { fprintf(stderr, "%s is shorter than %lld bytes\n", path, expect_bytes); exit(2); }
    fclose(f);
    return buf;
}

static mugd_region* find_region(const char* name) {
    for (int i = 0; i < n_regions; ++i)
        if (strcmp(names[i], name) == 0) return &regions[i];
    fprintf(stderr, "unknown region %s\n", name);
    exit(2);
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <bundle dir>\n", argv[0]); return 2; }
    const char* dir = argv[1];
    char path[1024], line[2048];
    snprintf(path, sizeof(path), "%s/manifest.txt", dir);
    FILE* mf = fopen(path, "r");
    if (!mf) { fprintf(stderr, "cannot open %s\n", path); return 2; }

    mugd_handle* h = NULL;
    MK(mugd_create(0, &h));
    MK(mugd_set_gemm_impl(h, MUGD_GEMM_TC));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int bad = 0;
    float sample_ms = 0.f;
    int sample_steps = 0;

    while (fgets(line, sizeof(line), mf)) {
        char a[64], b[256], c[256], d[256];
        long long nbytes;
        if (line[0] == '#' || sscanf(line, "%63s", a) != 1) continue;
        if (strcmp(a, "region") == 0) {
            if (sscanf(line, "%*s %47s %lld %255s %255s", names[n_regions], &nbytes, b, c) != 4 || n_regions >= MAXR) { fprintf(stderr, "bad line: %s", line); return 2; }
            mugd_region* r = &regions[n_regions];
            r->name = names[n_regions];
            r->bytes = nbytes;
            CK(cudaMalloc(&r->base, (size_t)nbytes));
            CK(cudaMemset(r->base, 0, (size_t)nbytes));
            if (strcmp(b, "file") == 0) {
                void* buf = read_file(dir, c, nbytes);
                CK(cudaMemcpy(r->base, buf, (size_t)nbytes, cudaMemcpyHostToDevice));
                free(buf);
            }
            ++n_regions;
        } else if (strcmp(a, "plan") == 0) {
            if (sscanf(line, "%*s %255s %255s", b, c) != 2) { fprintf(stderr, "bad line: %s", line); return 2; }
            snprintf(path, sizeof(path), "%s/%s", dir, b);
            mugd_plan* p = NULL;
            MK(mugd_plan_load(h, path, regions, n_regions, &p));
            if (strcmp(c, "graph") == 0) {
                MK(mugd_plan_run(p, st));                  /* warm-up outside capture (lazy module load) */
                MK(mugd_plan_capture(p, st));
                MK(mugd_plan_replay(p, 1, st));
            } else {
                MK(mugd_plan_run(p, st));
            }
            CK(cudaStreamSynchronize(st));
            printf("ran %-12s (%s, %d launches)\n", b, c, mugd_plan_launch_count(p));
            mugd_plan_destroy(p);
        } else if (strcmp(a, "sample") == 0) {
            int steps = 0;
            if (sscanf(line, "%*s %255s %255s %d", b, c, &steps) != 3) { fprintf(stderr, "bad line: %s", line); return 2; }
            mugd_plan *pe = NULL, *pt = NULL;
            snprintf(path, sizeof(path), "%s/%s", dir, b);
            MK(mugd_plan_load(h, path, regions, n_regions, &pe));
            snprintf(path, sizeof(path), "%s/%s", dir, c);
            MK(mugd_plan_load(h, path, regions, n_regions, &pt));
            const mugd_op* tail = NULL;
            int32_t n_tail = 0;
            MK(mugd_plan_ops(pt, &tail, &n_tail));
            /* capture needs one eager pass first; that pass must not disturb the request, so save / restore the latent and counters:
             * here simply: run the eager warm-up BEFORE loadx would be wrong, so warm up on a copy of the state */
            mugd_region* arena = find_region("arena");
            void* snap = NULL;
            CK(cudaMalloc(&snap, (size_t)arena->bytes));
            CK(cudaMemcpy(snap, arena->base, (size_t)arena->bytes, cudaMemcpyDeviceToDevice));
            MK(mugd_plan_run(pe, st));
            MK(mugd_plan_capture(pe, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpy(arena->base, snap, (size_t)arena->bytes, cudaMemcpyDeviceToDevice));
            CK(cudaFree(snap));
            CK(cudaEventRecord(e0, st));
            MK(mugd_sample(pe, tail, n_tail, steps, st));   /* the whole DDIM loop: one call, no synchronisation inside */
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaEventElapsedTime(&sample_ms, e0, e1));
            sample_steps = steps;
            printf("sampled %d DDIM steps in %.3f ms (%.1f steps/s, %d launches per evaluation)\n", steps, sample_ms, 1000.0 * steps / sample_ms,
                   mugd_plan_launch_count(pe));
            mugd_plan_destroy(pe);
            mugd_plan_destroy(pt);
        } else if (strcmp(a, "expect") == 0) {
            if (sscanf(line, "%*s %63s %lld %255s", d, &nbytes, b) != 3) { fprintf(stderr, "bad line: %s", line); return 2; }
            mugd_region* r = find_region(d);
            float* got = (float*)malloc((size_t)nbytes);
            float* want = (float*)read_file(dir, b, nbytes);
            CK(cudaMemcpy(got, r->base, (size_t)nbytes, cudaMemcpyDeviceToHost));
            double maxabs = 0.0, maxerr = 0.0;
            long long n = nbytes / 4, nonfinite = 0;
            for (long long i = 0; i < n; ++i) {
                if (!isfinite(got[i])) ++nonfinite;
                if (fabs(want[i]) > maxabs) maxabs = fabs(want[i]);
                if (fabs((double)got[i] - want[i]) > maxerr) maxerr = fabs((double)got[i] - want[i]);
            }
            const double rel = maxerr / (maxabs > 0 ? maxabs : 1.0);
            /* same kernels, same plans, same data: equal up to the summation order of the fp64 row-moment atomics */
            const int ok = nonfinite == 0 && rel < 1e-5;
            printf("%-12s %lld values, max |x| %.4f, max abs diff to the Python run %.3e (rel %.2e)  %s\n", d, n, maxabs, maxerr, rel, ok ? "OK" : "MISMATCH");
            if (!ok) bad = 1;
            free(got);
            free(want);
        }
    }
    fclose(mf);
    (void)sample_steps;
    mugd_destroy(h);
    return bad;
}
