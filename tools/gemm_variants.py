"""A/B harness for the (T) GEMM tile variants (QCP2_T_VARIANT): times the segmented-K DMMA
GEMM through the resident (T) plan on config 5 (o=40, v=400), one process per variant."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import torch
    import qcp2_b200
    ctx = qcp2_b200.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_pointer_mode(qcp2_b200.DEVICE)
    o, v = 40, 400
    d = qcp2_b200.synth_t_inputs_device(ctx, o, v)
    plan = qcp2_b200.DeviceTriples(ctx, *(d[k] for k in ("T1", "T2", "oovv", "vovv", "ovoo", "eps_o", "eps_v")),
                                   qcp2_b200.T_SYMMETRIC)
    fpt = 2.0 * v * (v * (v - 1) // 2) * 3 * (v + o)
    plan.run(0, 1, 20)
    res = []
    for rep in range(3):
        e = plan.run(100 * rep, 1, 60)
        tm = plan.timings()
        res.append(dict(gemm_tflops=60 * fpt / tm["gemm_seconds"] / 1e12, total_tflops=60 * fpt / tm["total_seconds"] / 1e12,
                        e=e))
    print(json.dumps(dict(variant=os.environ.get("QCP2_T_GEMM", "tma") + ":" + os.environ.get("QCP2_T_VARIANT", "-"), runs=res)), flush=True)
    ctx.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for var in sys.argv[1:] or ["tma", "0", "1", "2"]:
            env = dict(os.environ, QCP2_T_VARIANT=var) if var != "tma" else dict(os.environ)
            if var != "tma":
                env["QCP2_T_GEMM"] = "cpasync"
            subprocess.call([sys.executable, os.path.abspath(__file__), "child"], env=env)
