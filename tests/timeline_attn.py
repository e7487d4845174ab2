"""Dump and summarise the per-block clock stamps of CTA 0 (perf experiment; needs the timeline build of the library):
    sh vorta_b200/csrc/build_variant.sh timeline -DVB_TIMELINE
    VB_LIB_PATH=vorta_b200/lib/exp/libvb_timeline.so python tests/timeline_attn.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vorta_b200 import ops

S, H = 16384, 37
plan = ops.Plan((1, 1, S), (1, 1, S), (1, 1, 1), (1, 1, 2), 0.5)
q, k, v = (torch.randn((1, S, H, 128), device="cuda").bfloat16().transpose(1, 2) for _ in range(3))
dbg = torch.zeros(4 * 64 * 8 * 2 + 70000, device="cuda")
for _ in range(3):
    ops.routed_attention(plan, q, k, v, branch=[0] * H, debug=dbg)
torch.cuda.synchronize()
t = dbg[:4 * 64 * 8 * 2].view(torch.int64).view(4, 64, 8).cpu()
base = t[0, 8, 1].item()
print("softmax tile t: [wait_start, s_ready, ld_done, exp_done, st_done, arrived]; mma t: [p_seen, k_ready, issued]")
for j in range(8, 20):
    for who in (0, 1):
        r = (t[who, j, :6] - base).tolist()
        print(f"j={j} softmax{who}: {r}  | wait {r[1]-r[0]:5d} ld {r[2]-r[1]:5d} exp {r[3]-r[2]:5d} st {r[4]-r[3]:5d}")
    for who in (2, 3):
        r = (t[who, j, :3] - base).tolist()
        print(f"j={j} mma  t{who-2}: {r}  | issue {r[2]-r[1]:5d}")
per = (t[0, 40, 1] - t[0, 8, 1]).item() / 32
print("avg period per pair-block (cycles):", per)
for who in (0, 1):
    d = t[who, 8:40]
    print(f"tile{who}: wait {float((d[:,1]-d[:,0]).float().mean()):.0f} ld {float((d[:,2]-d[:,1]).float().mean()):.0f} "
          f"compute {float((d[:,3]-d[:,2]).float().mean()):.0f} st_wait {float((d[:,4]-d[:,3]).float().mean()):.0f} "
          f"s_ready->arrived {float((d[:,5]-d[:,1]).float().mean()):.0f}")
for who in (0, 1):
    d = t[who, 8:40]
    print(f"tile{who}: ld_done->max+pair_sync {float((d[:,6]-d[:,2]).float().mean()):.0f}  ->first32 exps {float((d[:,7]-d[:,6]).float().mean()):.0f}"
          f"  ->all exps {float((d[:,3]-d[:,7]).float().mean()):.0f}")
m = t[2:4, 8:40]
print("mma: p_seen->issued", float((m[:, :, 2] - m[:, :, 0]).float().mean()))
# latency from softmax arrive to MMA seeing it, and from MMA issue end to S ready of next block
arr = t[0, 8:39, 5]; seen = t[2, 8:39, 0]
print("arrive->p_seen (tile0):", float((seen - arr).float().mean()))
issued = t[2, 8:39, 2]; nxt = t[0, 9:40, 1]
print("issued->next s_ready (tile0):", float((nxt - issued).float().mean()))
