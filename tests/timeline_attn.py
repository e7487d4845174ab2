"""Per-block clock stamps of CTA 0 of the attention kernel (perf experiment; needs the timeline build of the library):
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
print("softmax t: wait_start s_ready ld_done max_sync half_exps exp_done st_done arrived | mma t: qk_issue qk_issued pv0_issue pv1_issued")
for j in range(8, 16):
    for who in (0, 1):
        r = (t[who, j] - base).tolist()
        print(f"j={j} softmax{who}: wait {r[0]:6d} s_ready {r[1]:6d} ld {r[2]:6d} max {r[6]:6d} half {r[7]:6d} exp {r[3]:6d} st {r[4]:6d} arr {r[5]:6d}")
    for who in (2, 3):
        r = (t[who, j] - base).tolist()
        print(f"j={j} mma t{who-2}: qk_issue {r[0]:6d} qk_issued {r[3]:6d} pv0 {r[1]:6d} pv1_issued {r[2]:6d}")
per = (t[0, 40, 1] - t[0, 8, 1]).item() / 32
print("avg period per pair-block (cycles):", per)
for who in (0, 1):
    d = t[who, 8:40].float()
    print(f"tile{who}: wait {float((d[:,1]-d[:,0]).mean()):.0f} ld {float((d[:,2]-d[:,1]).mean()):.0f} max+sync {float((d[:,6]-d[:,2]).mean()):.0f} "
          f"exps {float((d[:,3]-d[:,6]).mean()):.0f} st {float((d[:,4]-d[:,3]).mean()):.0f} "
          f"s_ready->arrived {float((d[:,5]-d[:,1]).mean()):.0f}")
