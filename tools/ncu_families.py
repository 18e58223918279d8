"""Run one representative op of every kernel family of the real launch plan (largest of its kind) between
cudaProfilerStart/Stop, so that `ncu --profile-from-start off --set full` captures ~15 launches instead of a whole eval.
usage (GPU box): ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/families \
                     python tools/ncu_families.py [--B 4] [--L 512]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from mug_diffusion_b200 import lib as L_, synth  # noqa: E402
from mug_diffusion_b200.config import ModelConfig  # noqa: E402
from mug_diffusion_b200.engine import OpList  # noqa: E402
from mug_diffusion_b200.sampler import MugDiffusionB200  # noqa: E402
from profile_ops import signature  # noqa: E402


def work(op):
    sig = signature(op)
    if sig[0] == "gemm":
        return 2.0 * sig[1] * sig[2] * sig[3] * sig[4]
    if sig[0] == "attention":
        return 4.0 * sig[1] * sig[2] * sig[3] * sig[4] * sig[5]
    if sig[0] == "groupnorm":
        return sig[1] * sig[2] * sig[3]
    if sig[0] == "layernorm":
        return sig[1] * sig[2]
    if sig[0] == "s4conv":
        return sig[1] * sig[2] * sig[2] * sig[3]
    return 1.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--L", type=int, default=512)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    model = MugDiffusionB200(synth.synthetic_state_dict(a.L), ModelConfig(), z_length=a.L, device=dev)
    eng = model.engine
    sess = eng.session(2 * a.B, a.L, per_sample_t=False)
    sess.eval(graph=False)                      # every buffer holds real activations
    torch.cuda.synchronize()
    arr, n = sess.plan._arr, sess.plan.n_ops
    best = {}
    for i in range(n):
        sig = signature(arr[i])
        key = sig[0]
        if key == "gemm":                      # one per flavour: conv3, 1x1, gated, down/up, and the largest split-K case
            g = arr[i].u.gemm
            key = f"gemm:{'conv' + str(g.taps) if g.taps > 1 else '1x1'}:{'gate' + str(g.gate) if g.gate else 'plain'}"
        elif key == "attention":
            key = "attention:self" if sig[3] == sig[4] else "attention:cross"
        if key not in best or work(arr[i]) > work(arr[best[key]]):
            best[key] = i
    picked = sorted(best.items(), key=lambda kv: kv[1])
    for key, i in picked:
        print(f"{key:28s} op {i:3d} {signature(arr[i])}", flush=True)
    torch.cuda.profiler.start()
    for key, i in picked:
        sub = OpList()
        sub.ops.append(arr[i])
        eng.run_ops(sub)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
