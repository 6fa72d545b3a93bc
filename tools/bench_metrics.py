"""Time the device metric kernels (ug_depth_metrics / ug_normal_metrics) on one 25x384x512 clip against the
oracle's CPU restatement of the reference functions (1.4 s + 0.34 s per clip there).
Prints one JSON line; `python tools/bench_metrics.py [--out gpurun_out/metrics_bench.json]`."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    from oracle import metrics as OM
    from unigeo_b200.metrics import _default_engine
    eng = _default_engine()
    dev = eng.device
    shape = (25, 384, 512)
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(shape, generator=g) * 10 - 0.5
    pred = 0.4 * gt.clamp(min=0.1) + 1.0 + 0.1 * torch.randn(shape, generator=g)
    mask = torch.rand(shape, generator=g) > 0.25
    pn = torch.nn.functional.normalize(torch.randn(shape + (3,), generator=g), dim=-1)
    gn = torch.nn.functional.normalize(pn + 0.3 * torch.randn(shape + (3,), generator=g), dim=-1)
    d = [t.to(dev) for t in (pred, gt, mask, pn, gn)]
    n = pred.numel()

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e3

    res = {"clip": list(shape), "unit": "ms per clip (host wall clock around the synchronous C-ABI call)"}
    res["depth_device_resident_ms"] = timed(lambda: eng.depth_metrics(d[0], d[1], d[2]), a.iters)
    res["normal_device_resident_ms"] = timed(lambda: eng.normal_metrics(d[3], d[4], d[2]), a.iters)
    res["depth_host_inputs_ms"] = timed(lambda: eng.depth_metrics(pred, gt, mask), 3)
    res["normal_host_inputs_ms"] = timed(lambda: eng.normal_metrics(pn, gn, mask), 3)
    # algorithmic bytes: depth 2 passes x (pred + gt) + mask; normals (pred + gt) x 12 B + mask + err write + 4 select reads
    res["depth_gbps"] = (16.0 + 1.0) * n / (res["depth_device_resident_ms"] * 1e-3) / 1e9
    res["normal_gbps"] = (24.0 + 1.0 + 4.0 + 16.0) * n / (res["normal_device_resident_ms"] * 1e-3) / 1e9
    t0 = time.perf_counter()
    rd = OM.depth_evaluation(pred, gt, mask)
    res["depth_cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    rn = OM.normal_evaluation(pn, gn, mask)
    res["normal_cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    res["cpu_threads"] = torch.get_num_threads()
    vd, _ = eng.depth_metrics(d[0], d[1], d[2])
    vn = eng.normal_metrics(d[3], d[4], d[2])
    res["max_abs_delta_depth"] = max(abs(x - float(y)) for x, y in zip(vd[:8], list(rd.values())[:8]))
    res["max_abs_delta_normal"] = max(abs(x - float(y)) for x, y in zip(vn, rn.values()))
    res["valid_pixels_equal"] = int(vd[8]) == int(rd["valid_pixels"])
    line = json.dumps(res)
    print(line)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        open(a.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
