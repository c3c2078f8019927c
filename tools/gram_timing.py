"""Per-warp cycle attribution of the fused lift+Gram kernel (development build: NK_EXTRA_NVCC_FLAGS=-DNK_GRAM_TIMING)."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from nys_koop_lqr_b200.engine import Engine

NAMES = ["total", "wait_item", "pack", "wait_first_slab", "mainloop_lift", "mainloop_syrk", "flush_pending", "epilogue_lift",
         "epilogue_syrk", "n_lift", "n_syrk", "midloop_wait_cycles", "midloop_waits", "lift_kernel_function_loop", "lift_publish_fences"]


def main(n=200000, m=4096, d=192, p=6, chunk=512):
    eng = Engine.get()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    Xa = torch.randn(n, d + p, dtype=torch.float64, device="cuda", generator=g)
    Y = torch.tanh(Xa[:, :d] * 0.5)
    Z = Y[:m].contiguous()
    il = torch.full((d,), 0.1, dtype=torch.float64, device="cuda")
    for rep in range(2):
        eng.gram_begin(Z, il, 0, p, chunk)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.gram_update(Xa, Y); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    cnt = eng.sm_count() * 8 * 16
    buf = (C.c_longlong * cnt)()
    fn = eng.lib.nk_debug_gram_timing
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]
    assert fn(eng.h, buf, cnt) == 0
    t = np.array(buf, dtype=np.float64).reshape(eng.sm_count(), 8, 16)
    tot = t[:, :, 0].mean()
    out = {"ms": ms, "cycles_total_mean": tot, "mhz_implied": tot / ms * 1e-3}
    for i, nm in enumerate(NAMES):
        if i == 0:
            continue
        out[nm] = float(t[:, :, i].mean()) if nm.startswith("n_") or nm == "midloop_waits" else float(t[:, :, i].mean() / tot)
    # ideal main-loop cycles: DMMAs per warp x 16 clk x 2 warps per sub-partition
    KLS = (d + 2 + 15) // 16
    ideal = (t[:, :, 9] * KLS + t[:, :, 10] * (chunk // 16)) * 128 * 32
    out["cycles_per_lift_epilogue"] = float((t[:, :, 7].sum() / max(t[:, :, 9].sum(), 1)))
    out["cycles_per_lift_function_loop"] = float((t[:, :, 13].sum() / max(t[:, :, 9].sum(), 1)))
    out["cycles_per_lift_publish"] = float((t[:, :, 14].sum() / max(t[:, :, 9].sum(), 1)))
    out["cycles_per_syrk_epilogue"] = float((t[:, :, 8].sum() / max(t[:, :, 10].sum(), 1)))
    out["cycles_per_syrk_flush"] = float((t[:, :, 6].sum() / max((t[:, :, 9] + t[:, :, 10]).sum(), 1)))
    out["cycles_per_syrk_mainloop"] = float((t[:, :, 5].sum() / max(t[:, :, 10].sum(), 1)))
    out["cycles_per_lift_mainloop"] = float((t[:, :, 4].sum() / max(t[:, :, 9].sum(), 1)))
    out["mainloop_ideal_frac_of_total"] = float((ideal / t[:, :, 0]).mean())
    out["mainloop_measured_frac_of_total"] = out["mainloop_lift"] + out["mainloop_syrk"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
