#!/usr/bin/env python
"""bench.py -- radial-loop benchmark of magic_b200 (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of MagIC's radial-loop hot path over all radial levels of the workload:
    transp_lm2r (flow, s, field containers) -> radial loop (SHT synthesis, get_nl, SHT analysis, get_td)
    -> transp_r2lm (dflowdt, dsdt, dbdt containers)
i.e. `rLoop_counter + comm_counter` of the reference (step_time.f90:491-542,1016,1149).  Radial levels are
sharded over the N ranks with getBlocks (parallel.f90:75-92); the two transposes are all-to-alls over NVLink.

`value` counts the Legendre flops the REFERENCE spends on the step (U * 2*n_theta*lm_max per level, U = 36
for the MHD set: its vector transforms sum against Plm and dPlm) divided by the max-over-ranks device time of a step, so it
is comparable with the CPU arm.  The library itself executes fewer flops for the same result (dPlm is a 3-point
combination of Plm, so a vector component is one pass: 22 units for the MHD set); `roofline` is computed from the
EXECUTED flops and says so.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "radial-loop SHT GFLOP/s (FP64, reference-equivalent Legendre flops / radial-loop step time incl. r<->LM transposes)"
UNIT = "GFLOP/s"
FP64_PEAK_TFLOPS = 37.0  # DMMA.8x8x4 probe, profiles/fp64_peak_r01.json (MEASURED_PEAKS.json has no FP64 entry)
HBM_FALLBACK_GBS = 6546.6  # MEASURED_PEAKS.json of this pool (used when the driver-written file is absent)
DEFAULT_WORKLOAD = "dynamo_l1023"
# per physics: reference units per bulk level, grid fields synthesised / analysed per level (= FFTs),
# `module sht` calls per level (rIter.f90:466-712)
PHYS = {"mhd": dict(units=36, n_in=13, n_out=9, calls=8), "anel": dict(units=29, n_in=12, n_out=7, calls=8),
        "hydro": dict(units=21, n_in=7, n_out=6, calls=5)}
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02", "gemm_traffic.json")  # written from the ncu --set full capture


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="magic_b200", choices=["magic_b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--level-chunk", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-f1", default="on", choices=["on", "off"],
                    help="also time the end-to-end call with the LM-side prologue / epilogue on the device (SURVEY 8(f)1); e2e = the faster")
    ap.add_argument("--overlap", default="auto", choices=["auto", "on", "off"],
                    help="transposes pipelined chunk-wise against the compute (magic_rloop_run_lm_dev); auto = on for N > 1")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-levels", type=int, default=0)
    return ap.parse_args()


def flops_per_level(gs):
    return PHYS[gs["physics"]]["units"] * 2.0 * gs["n_theta_max"] * gs["lm_max"]


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (MEASURED_PEAKS.json absent): the pool's measured copy bandwidth of round 1"


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the restated native path (oracle/) on the host cores.  Only this leg may touch oracle/.
def _oracle_setup(gs, n_levels, threads):
    from oracle.oracle import Oracle, Params as OParams
    from magic_b200.workload import config_l_R, config_params, make_fields, make_radial, seed_for
    o = Oracle(gs["l_max"], minc=gs["minc"], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"], fast=True,
               threads=threads)
    p = config_params(gs)
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    mid = gs["n_r_max"] // 2
    rad = make_radial(gs["n_r_max"], gs["l_max"], nRstart=mid, nRstop=mid + n_levels - 1, l_R=config_l_R(gs),
                      anel=(gs["physics"] == "anel"))
    fields = make_fields(gs["physics"], o.lm2l, o.lm2m, n_levels, seed_for(gs["config_id"], 0))
    return o, op, rad, fields


def cpu_reference_sample(gs, n_levels, threads, reps=1):
    """Times orc_radial_loop for n_levels bulk levels of the workload; returns (GFLOP/s, seconds per level)."""
    o, op, rad, fields = _oracle_setup(gs, n_levels, threads)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        o.radial_loop(op, rad, fields)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return flops_per_level(gs) * n_levels / best * 1e-9, best / n_levels


def run_reference(args, gs):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_lev = args.cpu_levels or (1 if gs["l_max"] >= 511 else 4)
    o, op, rad, fields = _oracle_setup(gs, n_lev, threads)
    for _ in range(args.warmup):
        o.radial_loop(op, rad, fields)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.radial_loop(op, rad, fields)
    dt = (time.perf_counter() - t0) / args.steps
    # scale the sample (n_lev levels) to a whole step (n_r_max levels): levels are independent
    ms_step = dt / n_lev * gs["n_r_max"] * 1e3
    value = flops_per_level(gs) * n_lev / dt * 1e-9
    sample = (f"{n_lev} bulk level(s) of {gs['n_r_max']} per timed step ({dt * 1e3:.0f} ms measured per step; ms_per_step is that "
              f"scaled to all {gs['n_r_max']} levels, levels are independent), radial loop only (no transposes), restated native SHT "
              "(not magic.exe: no Fortran compiler on the box)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "ms_per_timed_sample": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args, gs, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "transforms_per_s": transforms_per_s(gs, gs["n_r_max"], ms_step),
    }
    print(json.dumps(line), flush=True)


def config_dict(args, gs, chunk):
    return {"workload": args.workload, "l_max": gs["l_max"], "n_r_max": gs["n_r_max"], "n_theta": gs["n_theta_max"],
            "n_phi": gs["n_phi_max"], "lm_max": gs["lm_max"], "minc": gs["minc"], "fields": gs["physics"],
            "flags": gs.get("flags", {}), "l_var_l": bool(gs.get("l_var_l")),
            "units_per_level": PHYS[gs["physics"]]["units"], "level_chunk": chunk, "l2": "inputs_exceed_l2",
            "polar_eps": float(os.environ.get("MAGIC_POLAR_EPS", "1e-40")),
            "parallelism": f"r-slabs x{args.gpus} (getBlocks) + all-to-all transposes over NVLink" +
                           (", pipelined chunk-wise under the compute" if (args.overlap == "on" or (args.overlap == "auto" and args.gpus > 1)) else "")}


def transforms_per_s(gs, n_levels, ms):
    """Absolute transform rates of a step: grid fields transformed (one Legendre + FFT pass each way counts once) and
    `module sht` procedure calls (scal_to_spat, torpol_to_spat, torpol_to_curl_spat, spat_to_qst, ... rIter.f90:466-712)."""
    ph = PHYS[gs["physics"]]
    s = ms * 1e-3
    return {"field_transforms_per_s": (ph["n_in"] + ph["n_out"]) * n_levels / s, "sht_calls_per_s": ph["calls"] * n_levels / s,
            "levels_per_s": n_levels / s, "definition": f"per level: {ph['n_in']} spectral->grid + {ph['n_out']} grid->spectral field "
            f"transforms = {ph['calls']} module-sht calls"}


# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7:
                    continue
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def bit_digest(torch, dist, tensors, world):
    """Decomposition-independent digest of the LM-distributed outputs: the sum modulo 2^64 of the raw bit patterns of every
    double, all-reduced over the ranks (integer addition is associative, so the value does not depend on how levels and
    modes are split: identical at N = 1, 2, 4, 8 exactly when every output bit is)."""
    acc = None
    for t in tensors:
        if t is None:
            continue
        s = torch.view_as_real(t).contiguous().view(torch.int64).sum()
        acc = s if acc is None else acc + s
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return format(int(acc.item()) & 0xFFFFFFFFFFFFFFFF, "016x")


def run_magic(args, gs):
    import torch
    import torch.distributed as dist
    from magic_b200 import RadialLoop, Sht, Transposer
    from magic_b200.riter import OUT_NAMES
    from magic_b200.transpose import unique_id
    from magic_b200.workload import config_l_R, config_params, make_radial, seed_for

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; magic_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]

    physics = gs["physics"]
    ph = PHYS[physics]
    mag = physics == "mhd"
    p = config_params(gs)
    double_curl = bool(p.l_double_curl)
    sht = Sht(gs["l_max"], m_max=gs["m_max"], minc=gs["minc"], n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"],
              device_id=local_rank)
    n_r_max, lm_max = gs["n_r_max"], gs["lm_max"]
    tr = Transposer(sht, n_r_max, 5, rank=rank, n_procs=world, nccl_id=nccl_id)
    nr_loc, nlm_loc = tr.nr_loc, tr.nlm_loc
    ext = torch.cuda.ExternalStream(sht.stream, device=dev)

    def calloc(*shape):
        return torch.zeros(*shape, dtype=torch.complex128, device=dev)

    # R-distributed containers (fields.f90:211-268) and their LM-distributed images.  Every radial level is drawn from its
    # own seeded stream, so the global fields -- and therefore every output bit -- do not depend on the number of ranks.
    gen = torch.Generator(device=dev)
    lm2l = torch.from_numpy(sht.lm2l.astype(np.float64)).to(dev)
    lm2m = torch.from_numpy(sht.lm2m).to(dev)
    scale = 1.0 / (lm2l + 1.0)
    m0 = (lm2m != 0).to(torch.float64)

    def rand_container(nf, zero_l0, tag):
        a = torch.empty(nf, nr_loc, lm_max, dtype=torch.complex128, device=dev)
        v = torch.view_as_real(a)
        for i in range(nr_loc):
            gen.manual_seed(seed_for(gs["config_id"], 0) * 4096 + (tr.nRstart + i) * 8 + tag)
            for f in range(nf):
                v[f, i].normal_(generator=gen)
        v[..., 1].mul_(m0[None, None, :])
        v.mul_(scale[None, None, :, None])
        for f in range(nf):
            if zero_l0[f]:
                v[f, :, lm2l == 0, :] = 0.0
        return a

    nf_dflow = 4 if double_curl else 3  # dflowdt container: dwdt, dzdt, dpdt (, dVxVhLM), dt_fieldsLast.f90:125-214
    if gs.get("checkpoint"):   # a real saturated state instead of random spectra
        from magic_b200.workload import checkpoint_containers
        ck = checkpoint_containers(os.path.join(ROOT, gs["checkpoint"]), lm_max, n_r_max)
        mine = slice(tr.nRstart - 1, tr.nRstop)
        flow_R, s_R, field_R = (torch.from_numpy(np.ascontiguousarray(ck[k][:, mine])).to(dev) for k in ("flow", "s", "field"))
    else:
        flow_R = rand_container(5, [True] * 5, 0)       # w, dw, ddw, z, dz
        s_R = rand_container(2, [False, False], 1)      # s, ds
        field_R = rand_container(5, [True] * 5, 2) if mag else None   # b, db, ddb, aj, dj
    flow_LM, s_LM = calloc(5, n_r_max, nlm_loc), calloc(2, n_r_max, nlm_loc)
    field_LM = calloc(5, n_r_max, nlm_loc) if mag else None
    torch.cuda.synchronize()
    tr.transp_r2lm_dev_n(5, flow_R.data_ptr(), flow_LM.data_ptr())
    tr.transp_r2lm_dev_n(2, s_R.data_ptr(), s_LM.data_ptr())
    if mag:
        tr.transp_r2lm_dev_n(5, field_R.data_ptr(), field_LM.data_ptr())
    ext.synchronize()
    dflow_R, ds_R = calloc(nf_dflow, nr_loc, lm_max), calloc(2, nr_loc, lm_max)
    db_R = calloc(3, nr_loc, lm_max) if mag else None
    dflow_LM, ds_LM = calloc(nf_dflow, n_r_max, nlm_loc), calloc(2, n_r_max, nlm_loc)
    db_LM = calloc(3, n_r_max, nlm_loc) if mag else None
    dtr = torch.zeros(nr_loc, dtype=torch.float64, device=dev)
    dth = torch.zeros(nr_loc, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    rad = make_radial(n_r_max, gs["l_max"], nRstart=tr.nRstart, nRstop=tr.nRstop, l_R=config_l_R(gs), anel=(physics == "anel"))
    chunk = args.level_chunk
    rl = RadialLoop(sht, p, rad, level_chunk=chunk)

    fin = {"w": flow_R[0], "dw": flow_R[1], "ddw": flow_R[2], "z": flow_R[3], "dz": flow_R[4], "s": s_R[0]}
    fout = {"dwdt": dflow_R[0], "dzdt": dflow_R[1], "dsdt": ds_R[0], "dVSrLM": ds_R[1]}
    if double_curl:
        fout["dVxVhLM"] = dflow_R[3]
    else:
        fout["dpdt"] = dflow_R[2]
    if mag:
        fin.update({"b": field_R[0], "db": field_R[1], "ddb": field_R[2], "aj": field_R[3], "dj": field_R[4]})
        fout.update({"dbdt": db_R[0], "djdt": db_R[1], "dVxBhLM": db_R[2]})
    fin_p = {k: v.data_ptr() for k, v in fin.items()}
    fout_p = {k: v.data_ptr() for k, v in fout.items()}

    stage_acc = {}
    exposed_acc = [0.0, 0.0]
    tev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tacc = {"transp_lm2r": 0.0}
    overlap = args.overlap == "on" or (args.overlap == "auto" and world > 1)
    nsteps = [0]

    def step_overlapped():
        # one call: the all-to-alls of level chunk c+1 (in) and c-1 (out) run on a second stream under the compute of chunk c
        rl.run_lm_dev(tr, flow_LM.data_ptr(), s_LM.data_ptr(), field_LM.data_ptr() if mag else 0, dflow_LM.data_ptr(),
                      ds_LM.data_ptr(), db_LM.data_ptr() if mag else 0, dtr.data_ptr(), dth.data_ptr())
        for k, v in rl.last_timing().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        ex = rl.last_exposed()
        exposed_acc[0] += ex[0]
        exposed_acc[1] += ex[1]
        nsteps[0] += 1

    def step():
        if overlap:
            return step_overlapped()
        tev[0].record(ext)
        tr.transp_lm2r_dev_n(5, flow_LM.data_ptr(), flow_R.data_ptr())
        tr.transp_lm2r_dev_n(2, s_LM.data_ptr(), s_R.data_ptr())
        if mag:
            tr.transp_lm2r_dev_n(5, field_LM.data_ptr(), field_R.data_ptr())
        tev[1].record(ext)
        rl.radialLoop_dev(fin_p, fout_p, dtr.data_ptr(), dth.data_ptr())
        tev[2].record(ext)
        tr.transp_r2lm_dev_n(nf_dflow, dflow_R.data_ptr(), dflow_LM.data_ptr())
        tr.transp_r2lm_dev_n(2, ds_R.data_ptr(), ds_LM.data_ptr())
        if mag:
            tr.transp_r2lm_dev_n(3, db_R.data_ptr(), db_LM.data_ptr())
        tev[3].record(ext)
        for k, v in rl.last_timing().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        # (the radial loop synchronises per chunk, so the lm2r events of this step have completed)
        tev[1].synchronize()
        tacc["transp_lm2r"] += tev[0].elapsed_time(tev[1])
        nsteps[0] += 1

    def barrier():
        ext.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    stage_acc.clear()
    exposed_acc[0] = exposed_acc[1] = 0.0
    nsteps[0] = 0
    tacc["transp_lm2r"] = 0.0
    launches0 = sht.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record(ext)
        for _ in range(args.steps):
            step()
        e1.record(ext)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1) / args.steps
    launches = sht.launch_count() - launches0
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item())
    total_flops = flops_per_level(gs) * n_r_max
    value = total_flops / (ms_step * 1e-3) * 1e-9
    stages = {k: v / args.steps for k, v in stage_acc.items()}
    if overlap:
        # what the pipelined call could not hide: before the first kernel of the first chunk / after the last kernel of the last
        stages["transp_exposed_head"], stages["transp_exposed_tail"] = exposed_acc[0] / args.steps, exposed_acc[1] / args.steps
    else:
        stages["transp_lm2r"] = tacc["transp_lm2r"] / max(nsteps[0], 1)
        stages["transp_r2lm_plus_wait"] = ms - stages["transp_lm2r"] - stages["total"]
    leg_ms = stages["legendre_syn"] + stages["legendre_an"]
    units_ref, units_exec = rl.legendre_units()
    unit_flops = 2.0 * gs["n_theta_max"] * gs["lm_max"] * nr_loc
    leg_exec_tflops = units_exec * unit_flops / (leg_ms * 1e-3) * 1e-12 if leg_ms > 0 else 0.0
    leg_ref_tflops = units_ref * unit_flops / (leg_ms * 1e-3) * 1e-12 if leg_ms > 0 else 0.0
    digest = bit_digest(torch, dist, [dflow_LM, ds_LM, db_LM], world)
    csum = torch.view_as_real(dflow_LM).abs().sum()
    if world > 1:
        dist.all_reduce(csum, op=dist.ReduceOp.SUM)
    checksum = float(csum.item())

    # ---- HBM-bound stages: algorithmic bytes of this rank's levels over their measured time ---------------------------
    hbm_gbs, hbm_src = hbm_peak()
    plane = 8.0 * gs["n_theta_max"] * gs["n_phi_max"]            # one grid field of one level
    tm = 16.0 * gs["n_theta_max"] * gs["n_m_max"]                 # one (theta,m)-space field of one level
    spec = 16.0 * gs["lm_max"]
    n_src = len(fin)
    hbm_bytes = {"fft_c2r": ph["n_in"] * (tm + plane) * nr_loc, "get_nl": (ph["n_in"] + ph["n_out"]) * plane * nr_loc,
                 "fft_r2c": ph["n_out"] * (plane + tm) * nr_loc, "prep": (n_src + ph["n_in"]) * spec * nr_loc,
                 "get_td": (ph["n_out"] + 5 + len(fout)) * spec * nr_loc}
    hbm = {}
    for k, b in hbm_bytes.items():
        t = stages.get(k, 0.0)
        hbm[k] = {"bytes": b, "ms": t, "GBps": b / (t * 1e-3) * 1e-9 if t > 0 else None, "frac": b / (t * 1e-3) * 1e-9 / hbm_gbs if t > 0 else None}
    hbm_ms = sum(stages.get(k, 0.0) for k in hbm_bytes)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(TRAFFIC_FILE))
        if tj.get("workload") == args.workload and tj.get("level_chunk") == rl_chunk(rl, chunk):
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass

    # ---- end to end through the host-container C-ABI call (what a Fortran type_mpicuda + rIter_cuda_t pair binds): HOST
    #      LM-distributed containers in, HOST LM-distributed explicit terms out; H2D, both transposes and D2H inside -------
    e2e = None
    if not args.no_e2e:
        def to_host(t):
            if t is None:
                return None
            hbuf = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            hbuf.copy_(t)
            return hbuf
        h_in = {"flow": to_host(flow_LM), "s": to_host(s_LM), "field": to_host(field_LM)}
        h_out = {"dflowdt": to_host(dflow_LM), "dsdt": to_host(ds_LM), "dbdt": to_host(db_LM)}
        def local_bits(ts):   # this rank's share of the digest
            return sum(int(torch.view_as_real(t).contiguous().view(torch.int64).sum().item()) for t in ts if t is not None) & 0xFFFFFFFFFFFFFFFF
        ref_bits = local_bits(h_out.values())   # the device-path results
        del flow_LM, s_LM, field_LM, dflow_LM, ds_LM, db_LM, fin, fin_p, flow_R, s_R, field_R, fout, fout_p, dflow_R, ds_R, db_R
        torch.cuda.empty_cache()
        np_in = {k: v.numpy() for k, v in h_in.items() if v is not None}
        np_out = {k: v.numpy() for k, v in h_out.items() if v is not None}
        for v in np_out.values():
            v[...] = 0
        h_dtr, h_dth = np.zeros(nr_loc), np.zeros(nr_loc)
        # bytes that cross PCIe per step: the fields the loop reads (ds is not) / every explicit term
        h2d = sum(v[:(1 if k == "s" else v.shape[0])].nbytes for k, v in np_in.items())
        d2h = sum(v.nbytes for v in np_out.values()) + h_dtr.nbytes + h_dth.nbytes
        def time_e2e():
            for _ in range(max(1, min(args.warmup, 2))):
                rl.run_lm(tr, np_in, np_out, h_dtr, h_dth)
            bits = local_bits(h_out.values())
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(ext):
                e0.record(ext)
                for _ in range(args.steps):
                    rl.run_lm(tr, np_in, np_out, h_dtr, h_dth)
                e1.record(ext)
            barrier()
            wall = (time.perf_counter() - t0) / args.steps * 1e3
            ems = torch.tensor([max(e0.elapsed_time(e1) / args.steps, wall)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            return float(ems.item()), bits

        def record(ms_e2e, h2d_b, d2h_b, path, extra):
            r = {"value": total_flops / (ms_e2e * 1e-3) * 1e-9, "unit": UNIT, "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
                 "ms_per_step": ms_e2e, "bytes_are": "per rank", "path": path}
            r.update(extra)
            return r

        # (A) every container crosses PCIe: 11 field arrays up (ds is not read), 8 explicit terms down
        ms_a, bits_a = time_e2e()
        sm = torch.tensor([1 if bits_a == ref_bits else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(sm, op=dist.ReduceOp.MIN)
        variants = {"full_containers": record(
            ms_a, h2d, d2h, "magic_rloop_run_lm (host LM-distributed containers in, host LM-distributed explicit terms out: PCIe up, "
            "lm2r, radial loop, r2lm, PCIe down, pipelined level chunk by level chunk)", {"bit_identical_to_device_path": bool(sm.item())})}
        # (B) SURVEY 8(f)1 on the device (magic_rloop_lm_options): dw, ddw, dz, db, ddb, dj come from w, z, b, aj by the radial-matrix
        #     GEMM in the LM distribution, finish_explicit_assembly runs after the outbound transposes -- only w, z, s, b, aj go
        #     up and dVSrLM, dVxBhLM (dVxVhLM) stay down.  Same kernels and level count; the derivative inputs of the loop are then
        #     D w instead of independent random spectra, so the results are not comparable bit for bit (parity of this path:
        #     tests/test_lm_side_gpu.py).
        if args.e2e_f1 == "on":
            from magic_b200.workload import cheb_matrices
            D1, D2 = cheb_matrices(n_r_max)
            full = make_radial(n_r_max, gs["l_max"], l_R=config_l_R(gs), anel=gs["physics"] == "anel")
            rl.set_radial_matrices(D1, D2)
            rl.set_lm_radial(full["or2"], full["orho1"], np.zeros(n_r_max), full["l_R"])
            rl.lm_options(derivs_on_device=True, finish_on_device=True)
            fld = lambda v: v[0].nbytes
            h2d_b = 2 * fld(np_in["flow"]) + fld(np_in["s"]) + (2 * fld(np_in["field"]) if "field" in np_in else 0)
            d2h_b = 3 * fld(np_out["dflowdt"]) + fld(np_out["dsdt"]) + (2 * fld(np_out["dbdt"]) if "dbdt" in np_out else 0) \
                + h_dtr.nbytes + h_dth.nbytes
            ms_b, _ = time_e2e()
            rl.lm_options(derivs_on_device=False, finish_on_device=False)
            variants["derivatives_and_finish_on_device"] = record(
                ms_b, h2d_b, d2h_b, "magic_rloop_run_lm with magic_rloop_lm_options(1, 1): only w, z, s, b, aj cross PCIe upwards (radial "
                "derivatives by a radial-matrix GEMM on the device), finish_explicit_assembly on the device after the outbound "
                "transposes (dVSrLM, dVxBhLM stay on the device)", {"parity": "tests/test_lm_side_gpu.py"})
        best = min(variants, key=lambda k: variants[k]["ms_per_step"])
        e2e = dict(variants[best])
        e2e["variant"] = best
        e2e["variants"] = {k: {kk: vv for kk, vv in v.items() if kk in ("ms_per_step", "value", "h2d_bytes_per_step", "d2h_bytes_per_step")}
                           for k, v in variants.items()}

    # ---- CPU baseline beside it (rank 0, N=1 only) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n_lev = args.cpu_levels or (2 if gs["l_max"] >= 511 else 8)
        gf, sec_per_level = cpu_reference_sample(gs, n_lev, threads)
        cpu = {"value": gf, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_lev} bulk levels of {n_r_max}, radial loop only, restated native SHT (oracle/, -O3 -mavx2 -mfma, OpenMP)",
               "s_per_step_extrapolated": sec_per_level * n_r_max}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": ("checkpoint: the saturated dynamo of samples/boussBenchSat/checkpoint_end.start (" + gs["checkpoint"] + "), radial "
                     "derivatives by Chebyshev collocation") if gs.get("checkpoint") else
                    "synthetic: torch.randn, one seeded stream per radial level (20261017+1000*config, level, container), "
                    "Re,Im~N(0,1)/(l+1), Im(m=0)=0; random-init, no checkpoint",
            "config": config_dict(args, gs, rl_chunk(rl, chunk)),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "transforms_per_s": transforms_per_s(gs, n_r_max, ms_step),
            "roofline": {"bound": "tensor", "kernel": "legendre_gemm_kernel (FP64 DMMA.8x8x4)", "achieved": leg_exec_tflops,
                         "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": leg_exec_tflops / FP64_PEAK_TFLOPS,
                         "flops_basis": f"{units_exec:g} scalar-equivalent passes per level x 2 n_theta lm_max: the algorithmic flops of the "
                                        "Plm-only formulation this library executes (polar-skipped tiles still counted, so the tensor "
                                        "pipe itself is busy for about 0.82 of this figure)",
                         "achieved_reference_units": leg_ref_tflops,
                         "reference_units_note": f"the same launches credited with the reference's {units_ref:g} passes per level "
                                                 "(it sums against Plm and dPlm); may exceed the peak: the saving is algorithmic",
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "measured DMMA m8n8k4 loop, tools/fp64_peak.cu -> profiles/fp64_peak_r01.json "
                                        "(MEASURED_PEAKS.json carries no FP64 figure)",
                         "share_of_step": leg_ms / ms_step,
                         "hbm": {"peak": hbm_gbs, "unit": "GB/s", "peak_source": hbm_src, "share_of_step": hbm_ms / ms_step,
                                 "kernels": hbm,
                                 "bytes_basis": "algorithmic (SURVEY.md 8d): FFT = one (theta,m) field + one grid field per transform, "
                                                "get_nl = grid fields in + out, prep / get_td = spectral fields in + out"}},
            "cpu_baseline": cpu,
            "stages_ms": stages, "checksum": checksum,
            "parity_check": {"digest": digest, "what": "sum mod 2^64 of the bit patterns of all LM-distributed outputs over all ranks; "
                             "the inputs do not depend on N, so equal digests at N = 1, 2, 4, 8 mean bit-identical results"},
            "sht_tflops_overall": value * 1e-3,
        }
        print(json.dumps(line), flush=True)
    rl.finalize()
    tr.destroy_comm()
    sht.finalize_sht()
    if world > 1:
        dist.destroy_process_group()


def rl_chunk(rl, requested):
    """the level chunk actually in use (the library's auto rule when 0 was requested)"""
    try:
        return rl.level_chunk()
    except Exception:
        return requested or "auto"


def main():
    # rank 0 prints ONE JSON line on stdout.  NCCL's INFO log (NCCL_DEBUG=INFO on the driver's boxes) would go to stdout too:
    # point it at stderr instead of silencing it, so the communicator lines stay visible to whoever captures the run.
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    args = parse()
    from magic_b200.workload import config_sizes
    gs = config_sizes(args.workload)
    if args.impl == "reference":
        run_reference(args, gs)
    else:
        run_magic(args, gs)


if __name__ == "__main__":
    main()
