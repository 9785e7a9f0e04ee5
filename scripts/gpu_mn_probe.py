"""Bring-up probe for the MN-major operand forms of okge_gemm_tf32_nt (run on a B200 through gpurun).

Checks C = A B^T against an fp64 product of TF32-pre-rounded inputs for every combination of operand layouts,
optionally under the descriptor overrides OKGE_DBG_MN / OKGE_DBG_SWZ (see gemm_tf32.cu) to find the encoding
the hardware accepts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_knowledge_graph_embeddings_b200 import kernels as K  # noqa: E402


def tf32_round(x):
    u = x.view(torch.int32)
    u = (u + 0x1000) & ~0x1FFF
    return u.view(torch.float32)


def panels_of(x):
    """[rows, K] -> K-panels [ceil(K/32), rows, 32] (zero tail)."""
    rows, k = x.shape
    kp = (k + 31) // 32 * 32
    buf = torch.zeros((rows, kp), dtype=torch.float32, device=x.device)
    buf[:, :k] = x
    return K.Panels(buf.view(rows, kp // 32, 32).permute(1, 0, 2).contiguous(), rows, k)


def run(shapes):
    dev = torch.device("cuda")
    worst = {}
    for (M, N, Kd) in shapes:
        g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + Kd)
        a = tf32_round(torch.randn(M, Kd, generator=g)).to(dev)
        b = tf32_round(torch.randn(N, Kd, generator=g)).to(dev)
        ref = a.double() @ b.double().t()
        scale = ref.abs().max().item()
        aT, bT = a.t().contiguous(), b.t().contiguous()
        forms_a = {"row": a, "kpan": panels_of(a), "col": K.ColMajor(aT), "mnpan": panels_of(aT).T}
        forms_b = {"row": b, "kpan": panels_of(b), "col": K.ColMajor(bT), "mnpan": panels_of(bT).T}
        for na, fa in forms_a.items():
            for nb, fb in forms_b.items():
                if na in ("row", "kpan") and nb in ("row", "kpan") and not (na == "row" and nb == "row"):
                    continue
                try:
                    out = K.gemm_nt(fa, fb, splits=1)
                    torch.cuda.synchronize()
                    err = (out.double() - ref).abs().max().item() / scale
                except Exception as ex:  # noqa: BLE001
                    err = float("nan")
                    print("   EXC", na, nb, str(ex)[:200])
                key = f"{na}x{nb}"
                worst[key] = max(worst.get(key, 0.0), err) if err == err else float("nan")
    return worst


if __name__ == "__main__":
    shapes = [(70, 40, 9000), (512, 512, 4096), (1000, 200, 64), (129, 257, 100)]
    configs = [("default", None, None)]
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        configs += [("lt1 swapped lbo/sbo", "1,512,4096,1024", None),
                    ("lt2 + plain SW128 (8-row atoms)", "2,4096,1024,1024", "1"),
                    ("lt2 + plain SW128 swapped", "2,1024,4096,1024", "1"),
                    ("lt1 + plain SW128", "1,4096,512,1024", "1"),
                    ("lt2 + ATOM_32B", "2,4096,1024,1024", None)]
    for name, mn, swz in configs:
        for k, v in (("OKGE_DBG_MN", mn), ("OKGE_DBG_SWZ", swz)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        res = run(shapes)
        print(f"[{name}] " + "  ".join(f"{k}:{v:.1e}" for k, v in res.items()), flush=True)
