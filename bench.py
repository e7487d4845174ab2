#!/usr/bin/env python
"""bench.py — DiT denoise-step time of the VORTA-routed Wan 2.1 transformer on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # N > 1: launched by torch.distributed.run, one rank per GPU
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)

A "step" is one transformer forward (= one denoise step, no CFG doubling) over one synthetic latent video with
random-init weights of the named architecture.  One JSON line is printed by rank 0 (see README / DESIGN.md section 6
for every field).  Workloads (BASELINE.json configs):
    wan14 (default): Wan2.1-T2V-14B, 720p x 81 f -> 21x45x80 = 75,600 tokens, 40 heads, 40 blocks  [configs[2]]
    wan13          : Wan2.1-T2V-1.3B, 480p x 81 f -> 21x30x52 = 32,760 tokens, 12 heads, 30 blocks [configs[1]]
The default is the same for every N so that the driver's 1/2/4/8 scaling ratios compare like with like (12 heads
do not divide by 8); the wan13 numbers are reported next to it at N = 1 in "aux".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "wan14": dict(model="wan2.1-t2v-14b", latent=(21, 45, 80), tile=(3, 9, 16), window=(3, 3, 3),
                  lowres_window=(3, 3, 2), rate=0.5, text_tokens=512,
                  name="Wan2.1-T2V-14B 720p x81f (21x45x80 = 75,600 tokens, 40 heads, 40 blocks), one denoise step"),
    "wan13": dict(model="wan2.1-t2v-1.3b", latent=(21, 30, 52), tile=(3, 10, 4), window=(3, 3, 3),
                  lowres_window=(3, 3, 2), rate=0.5, text_tokens=512,
                  name="Wan2.1-T2V-1.3B 480p x81f (21x30x52 = 32,760 tokens, 12 heads, 30 blocks), one denoise step"),
}
# contract test only (tests/test_bench_contract.py): tiny grid, 4 heads, 2 blocks; never a reported number
WORKLOADS["tiny"] = dict(model="wan-contract-test", latent=(4, 6, 8), tile=(2, 3, 4), window=(3, 3, 3),
                         lowres_window=(2, 3, 2), rate=0.5, text_tokens=16,
                         name="CONTRACT TEST ONLY: 2-block 4-head Wan shell, 4x6x8 = 192 tokens")
WORKLOADS["hunyuan"] = dict(model="hunyuanvideo", latent=(33, 45, 80), tile=(3, 9, 16), window=(3, 3, 3),
                            lowres_window=(3, 3, 2), rate=0.5, text_tokens=256, text_valid=64,
                            name="HunyuanVideo 720p x129f (33x45x80 = 118,800 video tokens + 256 text (64 valid), 24 heads, "
                                 "20 dual + 40 single blocks), one denoise step")
TAU_SPARSE = 0.3          # the reference's inference default (scripts/wan/inference.py:75)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops"))), burst=float(p["bf16_tflops"]),
                    hbm=float(p["hbm_gbs"]), source="MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)")
    return dict(bf16=1400.0, burst=1590.0, hbm=6650.0, source="fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)")


def recorded_traffic(workload):
    """DRAM bytes (read + write) per attention launch of this workload from the committed `ncu --set full` capture
    (profiles/attn_traffic.json, written from the capture by profiles/summarize_full.py --traffic); None when no
    capture of this workload has been committed."""
    path = os.path.join(ROOT, "profiles", "attn_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f).get(workload)
        return (float(rec["dram_bytes_per_launch"]), rec["source"]) if rec else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, f"/tmp/vb_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=self.out,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.remove(self.path)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(power), samples=len(sm),
                    reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores, on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------------------
def cpu_sample_ms(wl, threads=None):
    """One routed-attention layer sample: ONE head per branch at the workload's full sequence length, fp32, all
    host threads.  Returns per-branch per-head milliseconds."""
    import torch
    from oracle import vorta_oracle as O
    if threads:
        torch.set_num_threads(threads)
    lat = wl["latent"]
    S = lat[0] * lat[1] * lat[2]
    g = torch.Generator().manual_seed(1234)
    q, k, v = (torch.randn((1, 1, S, 128), generator=g).to(torch.bfloat16).float() for _ in range(3))
    info = O.get_group_info(lat, wl["lowres_window"], wl["rate"])
    out = {}
    for e, name in ((0, "full"), (1, "coreset"), (2, "sliding")):
        t0 = time.perf_counter()
        O.routed_attention(q, k, v, info, lat, wl["window"], wl["tile"], branch=torch.tensor([e]))
        out[name] = (time.perf_counter() - t0) * 1e3
    return out


def branch_counts(branches):
    c = [0, 0, 0]
    for layer in branches:
        for e in layer:
            c[e] += 1
    return c


def cpu_step_ms(per_head_ms, counts):
    return per_head_ms["full"] * counts[0] + per_head_ms["coreset"] * counts[1] + per_head_ms["sliding"] * counts[2]


def model_dims(wl):
    """(heads, attention layers) of the workload's architecture."""
    if wl["model"] == "hunyuanvideo":
        from vorta_b200.dit import HUNYUAN_CONFIGS
        c = HUNYUAN_CONFIGS[wl["model"]]
        return c.heads, c.num_layers + c.num_single_layers
    from vorta_b200.dit import WAN_CONFIGS
    c = WAN_CONFIGS[wl["model"]]
    return c.heads, c.num_layers


def run_reference(args):
    import torch
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                   # rank 0 alone runs the CPU arm
    heads, layers = model_dims(wl)
    cores = torch.get_num_threads()
    # the same head-branch mix the GPU arm reports is unknown here without a GPU: price a uniform 1/3 mix, which
    # is what random-init routers produce in expectation; the GPU arm's line carries its exact counts
    total_heads = heads * layers
    counts = [total_heads / 3.0] * 3
    for _ in range(args.warmup):
        cpu_sample_ms(dict(wl, latent=(3, 9, 16), tile=(3, 9, 16), lowres_window=(3, 3, 2)))   # tiny warm-up
    steps = []
    last = None
    for _ in range(args.steps):
        last = cpu_sample_ms(wl)
        steps.append(cpu_step_ms(last, counts))
    value = statistics.mean(steps)
    sample = (f"{args.steps} x (one head per branch: full / coreset / sliding at S={wl['latent'][0] * wl['latent'][1] * wl['latent'][2]}, "
              f"fp32 torch CPU SDPA, oracle port of the reference path); step = per-head times x {heads} heads x "
              f"{layers} layers at a uniform 1/3 branch mix; ATTENTION ONLY (the CPU linears are not timed)")
    line = dict(impl="reference", metric="dit_denoise_step_ms", value=value, unit="ms", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=value, higher_is_better=False, scaling="strong",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=wl["name"], tile=wl["tile"], window=wl["window"],
                            coreset_window=wl["lowres_window"], reduction_rate=wl["rate"], tau_sparse=TAU_SPARSE,
                            parallelism=f"host cpu, {cores} threads",
                            routing="uniform 1/3 branch mix (expectation of random-init routers)",
                            heads_per_branch_all_layers=dict(full=counts[0], coreset=counts[1], sliding=counts[2])),
                cpu_baseline=dict(value=value, unit="ms", cores=cores, kind="port", sample=sample,
                                  per_head_ms=last),
                e2e=dict(value=value, unit="ms", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def build_model(wl, device):
    import torch
    if wl["model"] == "hunyuanvideo":
        from vorta_b200.dit import HunyuanDiT
        from vorta_b200.patch import modeling_hunyuan, prepare_hunyuan_self_attn_kwargs
        model = HunyuanDiT.build(wl["model"], device, torch.bfloat16, seed=0)
        modeling_hunyuan.apply_vorta_transformer(model, train_router=False, router_dtype=torch.float32)
        kw = prepare_hunyuan_self_attn_kwargs(
            dict(latent_shape=wl["latent"], window_size=wl["window"], tile_size=wl["tile"],
                 lowres_window_size=wl["lowres_window"], lowres_reduction_rate=wl["rate"]), device, tau_sparse=TAU_SPARSE)
        return model, kw
    from vorta_b200.dit import WanDiT
    from vorta_b200.patch import apply_vorta_transformer, prepare_wan_self_attn_kwargs
    model = WanDiT.build(wl["model"], device, torch.bfloat16, seed=0)
    apply_vorta_transformer(model, train_router=False, router_dtype=torch.float32)
    kw = prepare_wan_self_attn_kwargs(
        dict(latent_shape=wl["latent"], window_size=wl["window"], tile_size=wl["tile"],
             lowres_window_size=wl["lowres_window"], lowres_reduction_rate=wl["rate"]), device, tau_sparse=TAU_SPARSE)
    return model, kw


def host_inputs(wl, cfg):
    import torch
    g = torch.Generator().manual_seed(1234)
    T, H, W = wl["latent"]
    lat = torch.randn((1, cfg.in_channels, T, 2 * H, 2 * W), generator=g).to(torch.bfloat16).pin_memory()
    text_dim = cfg.text_embed_dim if hasattr(cfg, "text_embed_dim") else cfg.text_dim
    txt = torch.randn((1, wl["text_tokens"], text_dim), generator=g).to(torch.bfloat16).pin_memory()
    ts = torch.tensor([500.0]).pin_memory()
    return lat, txt, ts


def make_step(model, wl, device, kw):
    """Returns call(latents, timestep, text) -> model outputs for the workload's model family."""
    import torch
    if wl["model"] != "hunyuanvideo":
        return lambda lat, ts, txt, **extra: model(lat, ts, txt, self_attention_kwargs=kw, return_dict=False, **extra)
    g = torch.Generator().manual_seed(99)
    mask = torch.zeros((1, wl["text_tokens"]), dtype=torch.bool, device=device)
    mask[:, :wl["text_valid"]] = True
    pooled = torch.randn((1, model.config.pooled_projection_dim), generator=g).to(device, torch.bfloat16)
    guidance = torch.tensor([6000.0], device=device)
    return lambda lat, ts, txt, **extra: model(lat, ts, txt, mask, pooled, guidance, self_attention_kwargs=dict(kw),
                                               return_dict=False, **extra)


def time_steps(fn, steps, warmup, dist_on, profile=False):
    import torch
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    if profile:          # ncu --profile-from-start off: capture exactly one warmed-up step, outside the timing
        torch.cuda.cudart().cudaProfilerStart()
        fn()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    ms = e0.elapsed_time(e1) / steps
    if dist_on:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run_workload(wl, args, rank, world, device, with_e2e=True):
    import torch
    import torch.distributed as dist
    from vorta_b200 import ops
    dist_on = world > 1
    model, kw = build_model(wl, device)
    cfg = model.config
    lat_h, txt_h, ts_h = host_inputs(wl, cfg)
    lat_d, txt_d, ts_d = lat_h.to(device), txt_h.to(device), ts_h.to(device)
    out_h = torch.empty(lat_h.shape, dtype=torch.bfloat16).pin_memory()

    call = make_step(model, wl, device, kw)

    @torch.no_grad()
    def step_resident():
        return call(lat_d, ts_d, txt_d)[0]

    @torch.no_grad()
    def step_e2e():
        a = lat_h.to(device, non_blocking=True)
        b = txt_h.to(device, non_blocking=True)
        c = ts_h.to(device, non_blocking=True)
        out = call(a, c, b)[0]
        out_h.copy_(out, non_blocking=True)
        return out

    # routing mix of this run (one untimed forward)
    with torch.no_grad():
        scores = call(lat_d, ts_d, txt_d, return_routing_scores=True)[4]      # reference 5-tuple
    branches = [s[0].float().argmax(-1).tolist() for s in scores]
    counts = branch_counts(branches)

    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    ops.stats_reset()
    ops.timing_enable(True)
    ops.timing_collect()
    sampler.start()
    ms = time_steps(step_resident, args.steps, args.warmup, dist_on, profile=args.profile)
    clocks = sampler.stop()
    ops.timing_enable(False)
    kernel_ms, kernel_launches, kernel_flops = ops.timing_collect()
    # timing events were also recorded during warm-up: normalise per step over warmup + steps
    n_all = args.steps + args.warmup + (1 if args.profile else 0)
    attn_ms_step, attn_flops_step = kernel_ms / n_all, kernel_flops / n_all
    launches_step = ops.stats()[0] / n_all
    e2e_ms = time_steps(step_e2e, args.steps, 1, dist_on) if with_e2e else None
    # whole-job attention flops: ranks hold disjoint head chunks
    if dist_on:
        t = torch.tensor([attn_flops_step, attn_ms_step], device="cuda", dtype=torch.float64)
        fl = t.clone(); dist.all_reduce(fl, op=dist.ReduceOp.SUM)
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        job_flops, attn_ms_max = float(fl[0].item()), float(mx[1].item())
    else:
        job_flops, attn_ms_max = attn_flops_step, attn_ms_step
    del model
    torch.cuda.empty_cache()
    return dict(ms=ms, e2e_ms=e2e_ms, clocks=clocks, counts=counts, attn_ms_step=attn_ms_step,
                attn_flops_step=attn_flops_step, kernel_launches_step=kernel_launches / n_all,
                launches_step=launches_step, job_flops=job_flops, attn_ms_max=attn_ms_max,
                h2d=lat_h.numel() * 2 + txt_h.numel() * 2 + 4, d2h=out_h.numel() * 2, cfg=cfg)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from vorta_b200 import _lib as L
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: vorta_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    L.check(L.lib().vb_device_check())
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        from vorta_b200.ulysses import SP_STATE
        SP_STATE.setup_sp_group(world)
    wl = WORKLOADS[args.workload]
    r = run_workload(wl, args, rank, world, device)
    peaks = measured_peaks()
    cfg = r["cfg"]
    achieved = r["attn_flops_step"] / (r["attn_ms_step"] * 1e-3) / 1e12 if r["attn_ms_step"] > 0 else 0.0
    traffic, traffic_source = recorded_traffic(args.workload)
    line = dict(
        metric="dit_denoise_step_ms", value=r["ms"], unit="ms", n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=r["ms"], higher_is_better=False, scaling="strong", vs_baseline=None, dtype="bf16",
        data="synthetic (seeded N(0,1) latents / text embeddings, random-init weights)",
        config=dict(workload=wl["name"], tile=wl["tile"], window=wl["window"], coreset_window=wl["lowres_window"],
                    reduction_rate=wl["rate"], tau_sparse=TAU_SPARSE, parallelism=f"ulysses{world}" if world > 1 else "single",
                    routing="random-init routers, top-1 per head", heads_per_branch_all_layers=dict(
                        full=r["counts"][0], coreset=r["counts"][1], sliding=r["counts"][2]),
                    l2="per-step working set (weights + activations) is far larger than the 126 MB L2; no flush needed"),
        routed_attn_effective_tflops=r["job_flops"] / (r["attn_ms_max"] * 1e-3) / 1e12 if r["attn_ms_max"] > 0 else 0.0,
        attn_kernel_ms_per_step=r["attn_ms_step"],
        roofline=dict(bound="tensor", achieved=achieved, peak=peaks["bf16"], unit="TFLOP/s",
                      frac=achieved / peaks["bf16"], traffic=traffic, traffic_unit="bytes / launch (DRAM read + write)",
                      traffic_source=traffic_source, kernel="vb_attn_fwd_kernel",
                      frac_of_burst=achieved / peaks["burst"], frac_of_spec_2250=achieved / 2250.0,
                      launches_per_step=r["kernel_launches_step"], peak_source=peaks["source"]),
        clocks=r["clocks"],
        e2e=dict(value=r["e2e_ms"], unit="ms", h2d_bytes_per_step=r["h2d"], d2h_bytes_per_step=r["d2h"]),
        gpu_launches=int(round(r["launches_step"] * args.steps)),
    )
    if world == 1 and rank == 0:
        # CPU baseline: the oracle port on the host cores, bounded sample (one head per branch)
        if not args.no_cpu_baseline:
            per_head = cpu_sample_ms(wl)
            S = wl["latent"][0] * wl["latent"][1] * wl["latent"][2]
            line["cpu_baseline"] = dict(
                value=cpu_step_ms(per_head, r["counts"]), unit="ms", cores=torch.get_num_threads(), kind="port",
                per_head_ms=per_head,
                sample=(f"one head per branch (full / coreset / sliding) at S={S}, fp32 torch CPU SDPA (oracle port of the "
                        f"reference path), timed once; value = per-head ms x this run's head-branch counts over all "
                        f"{model_dims(wl)[1]} layers; ATTENTION ONLY, the CPU linears are not timed"))
        if args.workload == "wan14" and not args.no_aux:
            a = run_workload(WORKLOADS["wan13"], args, rank, world, device, with_e2e=True)
            ach = a["attn_flops_step"] / (a["attn_ms_step"] * 1e-3) / 1e12
            line["aux"] = dict(workload=WORKLOADS["wan13"]["name"], ms_per_step=a["ms"], e2e_ms=a["e2e_ms"],
                               heads_per_branch_all_layers=dict(full=a["counts"][0], coreset=a["counts"][1],
                                                                sliding=a["counts"][2]),
                               attn_kernel_ms_per_step=a["attn_ms_step"], attn_tflops=ach,
                               roofline_frac=ach / peaks["bf16"])
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vorta_b200", choices=["vorta_b200", "reference"])
    ap.add_argument("--workload", default="wan14", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--profile", action="store_true", help="bracket one warmed-up step with cudaProfilerStart/Stop")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
