"""Where does a tapgemm tile's time go?  Builds the library with -DUG_TAPGEMM_TRACE (clock64 stamps per warp role and
tile, compiled out of the product build -- its SASS is byte-identical with and without the macros), runs ONE launch of
the given shape and prints, averaged over CTAs and steady-state tiles (tile index >= 2), in SM cycles:

  period          tile-to-tile distance of the epilogue's end stamps (what the launch time is made of)
  producer        wait for the first free stage of a tile / first load issued -> last load issued
  mma             wait for a drained accumulator / wait for the first stage to land / first stage -> last commit
  epilogue g0|g1  wait for the accumulator / accumulator ready -> TMEM drained / drained -> tile's stores issued

    python tools/trace_tapgemm.py build                     # here (no GPU needed): compile the trace variant
    python tools/trace_tapgemm.py linear M K N [res|geglu]  # on the GPU box
    python tools/trace_tapgemm.py conv T H W C Cout
Development aid; nothing in the product or the tests loads the trace library."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# split1 / split2 / split3: stamp 1 of the epilogue moves to before the bias staging / after its barrier / just before
# the accumulator wait (isolates the parts of the tile prologue)
SPLIT = next((int(a[5:]) for a in sys.argv if a.startswith("split") and a[5:].isdigit()), 0)
TRACE_LIB = os.path.join(ROOT, "unigeo_b200", f"libunigeo_b200_trace_split{SPLIT}.so" if SPLIT else "libunigeo_b200_trace.so")
VARIANT = (f"trace_split{SPLIT}", ["UG_TAPGEMM_TRACE", f"UG_TRACE_SPLIT={SPLIT}"]) if SPLIT else ("trace", ["UG_TAPGEMM_TRACE"])
TILES, ROLES = 16, 4


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        return
    if sys.argv[1] == "build":
        from unigeo_b200.build import build_variant
        print(build_variant(*VARIANT))
        return
    if not os.path.exists(TRACE_LIB):
        from unigeo_b200.build import build_variant
        build_variant(*VARIANT)
    os.environ["UG_LIB"] = TRACE_LIB
    import ctypes as C
    import math

    import torch
    from unigeo_b200 import _lib, ops
    lib = _lib.load()
    fn = lib.ug_debug_tapgemm_trace
    fn.argtypes, fn.restype = [C.c_void_p], C.c_int
    dev = torch.device("cuda", 0)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).half()
    kind, dims = sys.argv[2 - 1], [int(v) for v in sys.argv[2:] if v.lstrip("-").isdigit()]
    if kind == "linear":
        M, K, N = dims[:3]
        x, W = rnd(M, K), rnd(N, K, sc=1 / math.sqrt(K))
        res = rnd(M, N) if "res" in sys.argv else None
        b = torch.zeros(N, device=dev)
        if "geglu" in sys.argv:                     # N = 2 x hidden, rows interleaved like ug_ctx_finalize does
            W, b = ops.geglu_interleave(W, b)
            run = lambda: ops.linear(x, W, bias=b, geglu=True)
        else:
            run = lambda: ops.linear(x, W, bias=b, res=res)
    elif kind == "conv":
        T, H, Wd, Cc, Co = dims[:5]
        x, W = rnd(T, H, Wd, Cc), rnd(9, Co, Cc, sc=1 / math.sqrt(9 * Cc))
        run = lambda: ops.conv3x3(x, W)
    else:
        raise SystemExit(__doc__)
    for _ in range(3):
        run()
    buf = torch.zeros((sms, ROLES, TILES, 4), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    fn(buf.data_ptr())
    run()
    torch.cuda.synchronize()
    fn(None)
    t = buf.cpu().double()
    used = t[:, 2, :, 3] > 0                                   # [unit, tile]: epilogue group 0 finished this tile
    n_units = int(used[:, 0].sum())
    tiles_per = used.sum(1)[used[:, 0]].float().mean().item()
    print(f"{kind} {dims}: {n_units} CTAs / CTA pairs stamped, {tiles_per:.1f} tiles each (first {TILES} recorded)")

    def avg(a, b_, lo=2):
        """mean over CTAs and tiles >= lo of (stamp a - stamp b); a, b_ = (role, slot)"""
        d = t[:, a[0], lo:, a[1]] - t[:, b_[0], lo:, b_[1]]
        ok = (t[:, a[0], lo:, a[1]] > 0) & (t[:, b_[0], lo:, b_[1]] > 0)
        return float(d[ok].mean()) if ok.any() else float("nan")

    end = t[:, 2, :, 3]
    ok = (end[:, 3:] > 0) & (end[:, 2:-1] > 0)
    period = float((end[:, 3:] - end[:, 2:-1])[ok].mean()) if ok.any() else float("nan")
    print(f"period                      {period:9.0f} clk")
    print(f"producer  stage wait        {avg((0, 1), (0, 0)):9.0f}   first -> last load issued {avg((0, 2), (0, 1)):9.0f}")
    print(f"mma       accumulator wait  {avg((1, 1), (1, 0)):9.0f}   first stage wait {avg((1, 2), (1, 1)):9.0f}   "
          f"first -> last stage {avg((1, 3), (1, 2)):9.0f}")
    for grp in (2, 3):
        print(f"epilogue g{grp - 2} accumulator wait {avg((grp, 1), (grp, 0)):9.0f}   ready -> drained "
              f"{avg((grp, 2), (grp, 1)):9.0f}   drained -> stores issued {avg((grp, 3), (grp, 2)):9.0f}")
    print(f"mma(i) issued -> epilogue g0 sees accumulator {avg((2, 1), (1, 3)):9.0f}")
    if "raw" in sys.argv:            # per-tile timeline of a few CTA pairs, clocks relative to the pair's first stamp
        for u in (0, n_units // 2):
            base = t[u][t[u] > 0].min()
            print(f"-- unit {u}: tile | producer wait0 first last | mma start accfree stage0 last | g0 start acc drained done | g1 ...")
            for i in range(TILES):
                if t[u, 2, i, 3] <= 0:
                    break
                row = [f"{i:2d} |"]
                for role, slots in ((0, 3), (1, 4), (2, 4), (3, 4)):
                    row += [f"{int(t[u, role, i, k] - base):7d}" if t[u, role, i, k] > 0 else "      -" for k in range(slots)] + ["|"]
                print(" ".join(row))


if __name__ == "__main__":
    main()
