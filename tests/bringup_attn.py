"""GPU bring-up diagnostics for the tcgen05 attention kernel (run with gpurun; not a pytest).

Prints, for a ladder of cases, the error of the raw score tile S = Q K^T (debug dump) and of the final output
against an fp32 SDPA reference computed on the same device, then small-grid branch checks against the oracle.

The raw-score dump is compiled only into bring-up builds of the library (it costs ~4 % in the hot loop):
    sh vorta_b200/csrc/build_variant.sh debug -DVB_DEBUG_DUMP
    VB_LIB_PATH=vorta_b200/lib/exp/libvb_debug.so python tests/bringup_attn.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vorta_oracle as O  # noqa: E402
from vorta_b200 import _lib as L  # noqa: E402
from vorta_b200 import ops  # noqa: E402


def ref_sdpa(q, k, v):
    return torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())


def report(name, out, ref):
    out, ref = out.float(), ref.float()
    cos = torch.nn.functional.cosine_similarity(out.flatten(), ref.flatten(), dim=0).item()
    mx = (out - ref).abs().max().item()
    print(f"{name:50s} cos={cos:.6f} max_abs={mx:.4e} ref_absmax={ref.abs().max().item():.3e}", flush=True)
    return cos, mx


def case_full(S, H=1, mode="rand", dbg=True, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    lat = (1, 1, S)
    plan = ops.Plan(lat, (1, 1, S), (1, 1, 1), (1, 1, 2), 0.5)
    shp = (1, S, H, 128)
    q = torch.randn(shp, device=dev).bfloat16().transpose(1, 2)
    k = torch.randn(shp, device=dev).bfloat16().transpose(1, 2)
    v = torch.randn(shp, device=dev).bfloat16().transpose(1, 2)
    if mode == "qzero":
        q = torch.zeros_like(q)
    if mode == "vkey":      # V[key, d] depends on the key only
        v = (torch.arange(S, device=dev).float()[None, None, :, None] % 64 / 64).expand(1, H, S, 128).bfloat16().contiguous()
    if mode == "vchan":     # V[key, d] depends on the channel only
        v = (torch.arange(128, device=dev).float()[None, None, None, :] / 128).expand(1, H, S, 128).bfloat16().contiguous()
    debug = torch.zeros(4 * 128 * 128 + 512, device=dev) if dbg else None
    out = ops.routed_attention(plan, q, k, v, branch=[0] * H, debug=debug)
    torch.cuda.synchronize()
    ref = ref_sdpa(q, k, v)
    report(f"full S={S} H={H} {mode}", out, ref)
    if dbg:
        s_dump = debug[:2 * 128 * 128].view(2, 128, 128)
        nq = min(S, 128)
        s_ref = q[0, 0, :nq].float() @ k[0, 0, :nq].float().T
        report("   raw S tile0 (q0:128 x k0:128)", s_dump[0, :nq, :nq], s_ref)
        if S >= 256:
            s_ref1 = q[0, 0, 128:256].float() @ k[0, 0, :128].float().T
            report("   raw S tile1 (q128:256 x k0:128)", s_dump[1], s_ref1)
    return out, ref


def main():
    print("device:", torch.cuda.get_device_name(0), "lib version", L.lib().vb_version(), flush=True)
    L.check(L.lib().vb_device_check())
    for mode in ("rand", "qzero", "vkey", "vchan"):
        case_full(128, 1, mode)
    case_full(256, 1, "rand")
    case_full(256, 2, "rand")
    case_full(384, 1, "rand")
    case_full(200, 1, "rand")     # tail masking
    case_full(1000, 2, "rand")
    case_full(4096, 4, "rand", dbg=False)
    case_full(8192, 2, "rand", dbg=False)

    # sliding tile + coreset on a small grid against the oracle (fp32 on CPU)
    torch.manual_seed(1)
    lat, tile, win, lw = (4, 8, 12), (2, 4, 4), (3, 3, 3), (2, 2, 2)
    S = lat[0] * lat[1] * lat[2]
    H = 3
    plan = ops.Plan(lat, tile, win, lw, 0.5)
    shp = (1, S, H, 128)
    q = torch.randn(shp).bfloat16().transpose(1, 2)
    k = torch.randn(shp).bfloat16().transpose(1, 2)
    v = torch.randn(shp).bfloat16().transpose(1, 2)
    info = O.get_group_info(lat, lw, 0.5)
    for e, name in ((0, "full"), (1, "coreset"), (2, "sliding")):
        out = ops.routed_attention(plan, q.cuda(), k.cuda(), v.cuda(), branch=[e] * H)
        torch.cuda.synchronize()
        ref = O.routed_attention(q.float(), k.float(), v.float(), info, lat, win, tile, branch=torch.tensor([e] * H))
        report(f"small grid {lat} branch {name}", out.cpu(), ref)
    out = ops.routed_attention(plan, q.cuda(), k.cuda(), v.cuda(), branch=[0, 1, 2])
    ref = O.routed_attention(q.float(), k.float(), v.float(), info, lat, win, tile, branch=torch.tensor([0, 1, 2]))
    report("small grid routed mix [0,1,2]", out.cpu(), ref)
    w = torch.softmax(torch.randn(1, H, 3), -1)
    out = ops.routed_attention(plan, q.cuda(), k.cuda(), v.cuda(), weights=w)
    ref = O.routed_attention(q.float(), k.float(), v.float(), info, lat, win, tile, weights=w)
    report("small grid blend", out.cpu(), ref)

    # coreset index tables
    un, po = ops.coreset_select(plan, q.cuda())
    un_ref, po_ref = O.match(q.double(), info)
    print("coreset tables equal (fp64 oracle):", bool((un.cpu() == un_ref).all()), bool((po.cpu() == po_ref).all()),
          flush=True)

    # quick timing: full attention 16k x 8 heads
    S, H = 16384, 8
    plan = ops.Plan((1, 1, S), (1, 1, S), (1, 1, 1), (1, 1, 2), 0.5)
    q = torch.randn((1, S, H, 128), device="cuda").bfloat16().transpose(1, 2)
    k = torch.randn_like(q)
    v = torch.randn_like(q)
    for _ in range(2):
        ops.routed_attention(plan, q, k, v, branch=[0] * H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        ops.routed_attention(plan, q, k, v, branch=[0] * H)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 4.0 * S * S * 128 * H
    print(f"full S={S} H={H}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
