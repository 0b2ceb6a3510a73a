#!/usr/bin/env python
"""bench.py -- radial-loop benchmark of magic_b200 (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of MagIC's radial-loop hot path over all radial levels of the workload:
    transp_lm2r (flow, s, field containers) -> radial loop (SHT synthesis, get_nl, SHT analysis, get_td)
    -> transp_r2lm (dflowdt, dsdt, dbdt containers)
i.e. `rLoop_counter + comm_counter` of the reference (step_time.f90:491-542,1016,1149).  Radial levels are
sharded over the N ranks with getBlocks (parallel.f90:75-92); the two transposes are NCCL all-to-alls.
`value` counts the ALGORITHMIC FP64 flops of the Legendre stage (U * 2*n_theta*lm_max per level,
U=36 for the MHD set) of all ranks divided by the max-over-ranks device time of a step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "radial-loop SHT GFLOP/s (FP64, algorithmic Legendre flops / radial-loop step time incl. r<->LM transposes)"
UNIT = "GFLOP/s"
FP64_PEAK_TFLOPS = 37.0  # DMMA.8x8x4 probe, profiles/fp64_peak_r01.json (MEASURED_PEAKS.json has no FP64 entry)
DEFAULT_WORKLOAD = "dynamo_l1023"
UNITS = {"mhd": 36, "anel": 29, "hydro": 21}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="magic_b200", choices=["magic_b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--level-chunk", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--overlap", default="auto", choices=["auto", "on", "off"],
                    help="transposes pipelined chunk-wise against the compute (magic_rloop_run_lm_dev); auto = on for N > 1")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-levels", type=int, default=0)
    return ap.parse_args()


def flops_per_level(gs):
    return UNITS[gs["physics"]] * 2.0 * gs["n_theta_max"] * gs["lm_max"]


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the restated native path (oracle/) on the host cores.  Only this leg may touch oracle/.
def cpu_reference_sample(gs, n_levels, threads, reps=1):
    """Times orc_radial_loop for n_levels bulk levels of the workload; returns (GFLOP/s, seconds per level)."""
    from oracle.oracle import Oracle, Params as OParams
    from magic_b200.workload import make_fields, make_params, make_radial, seed_for
    o = Oracle(gs["l_max"], minc=gs["minc"], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"], fast=True,
               threads=threads)
    p = make_params(gs["physics"], gs["n_r_max"])
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    mid = gs["n_r_max"] // 2
    rad = make_radial(gs["n_r_max"], gs["l_max"], nRstart=mid, nRstop=mid + n_levels - 1, anel=(gs["physics"] == "anel"))
    fields = make_fields(gs["physics"], o.lm2l, o.lm2m, n_levels, seed_for(gs["config_id"], 0))
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
    from oracle.oracle import Oracle, Params as OParams
    from magic_b200.workload import make_fields, make_params, make_radial, seed_for
    o = Oracle(gs["l_max"], minc=gs["minc"], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"], fast=True,
               threads=threads)
    p = make_params(gs["physics"], gs["n_r_max"])
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    mid = gs["n_r_max"] // 2
    rad = make_radial(gs["n_r_max"], gs["l_max"], nRstart=mid, nRstop=mid + n_lev - 1, anel=(gs["physics"] == "anel"))
    fields = make_fields(gs["physics"], o.lm2l, o.lm2m, n_lev, seed_for(gs["config_id"], 0))
    for _ in range(args.warmup):
        o.radial_loop(op, rad, fields)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.radial_loop(op, rad, fields)
    dt = (time.perf_counter() - t0) / args.steps
    # scale the sample (n_lev levels) to a whole step (n_r_max levels): levels are independent
    ms_step = dt / n_lev * gs["n_r_max"] * 1e3
    value = flops_per_level(gs) * n_lev / dt * 1e-9
    sample = f"{n_lev} bulk level(s) of {gs['n_r_max']} per step, radial loop only (no transposes), restated native SHT (not magic.exe)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(args, gs, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def config_dict(args, gs, chunk):
    return {"workload": args.workload, "l_max": gs["l_max"], "n_r_max": gs["n_r_max"], "n_theta": gs["n_theta_max"],
            "n_phi": gs["n_phi_max"], "lm_max": gs["lm_max"], "minc": gs["minc"], "fields": gs["physics"],
            "units_per_level": UNITS[gs["physics"]], "level_chunk": chunk, "l2": "inputs_exceed_l2",
            "polar_eps": float(os.environ.get("MAGIC_POLAR_EPS", "1e-40")),
            "parallelism": f"r-slabs x{args.gpus} (getBlocks) + NCCL all-to-all transposes" +
                           (", pipelined chunk-wise under the compute" if (args.overlap == "on" or (args.overlap == "auto" and args.gpus > 1)) else "")}


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


def run_magic(args, gs):
    import torch
    import torch.distributed as dist
    from magic_b200 import RadialLoop, Sht, Transposer
    from magic_b200.riter import OUT_NAMES
    from magic_b200.transpose import unique_id
    from magic_b200.workload import make_params, make_radial, seed_for

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
    if physics != "mhd":
        raise SystemExit("bench.py: the timed workloads use the MHD field set (north-star); pick an mhd workload")
    sht = Sht(gs["l_max"], m_max=gs["m_max"], minc=gs["minc"], n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"],
              device_id=local_rank)
    n_r_max, lm_max = gs["n_r_max"], gs["lm_max"]
    tr = Transposer(sht, n_r_max, 5, rank=rank, n_procs=world, nccl_id=nccl_id)
    nr_loc, nlm_loc = tr.nr_loc, tr.nlm_loc
    ext = torch.cuda.ExternalStream(sht.stream, device=dev)

    def calloc(*shape):
        return torch.zeros(*shape, dtype=torch.complex128, device=dev)

    # R-distributed containers (fields.f90:211-268) and their LM-distributed images
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed_for(gs["config_id"], rank))
    lm2l = torch.from_numpy(sht.lm2l.astype(np.float64)).to(dev)
    lm2m = torch.from_numpy(sht.lm2m).to(dev)
    scale = 1.0 / (lm2l + 1.0)

    def rand_container(nf, zero_l0):
        a = torch.empty(nf, nr_loc, lm_max, dtype=torch.complex128, device=dev)
        v = torch.view_as_real(a)
        for f in range(nf):
            v[f].normal_(generator=gen)
            v[f, :, :, 1].mul_((lm2m != 0).to(torch.float64))
            v[f].mul_(scale[None, :, None])
            if zero_l0[f]:
                v[f, :, lm2l == 0, :] = 0.0
        return a

    flow_R = rand_container(5, [True] * 5)       # w, dw, ddw, z, dz
    s_R = rand_container(2, [False, False])      # s, ds
    field_R = rand_container(5, [True] * 5)      # b, db, ddb, aj, dj
    flow_LM, s_LM, field_LM = calloc(5, n_r_max, nlm_loc), calloc(2, n_r_max, nlm_loc), calloc(5, n_r_max, nlm_loc)
    torch.cuda.synchronize()
    tr.transp_r2lm_dev_n(5, flow_R.data_ptr(), flow_LM.data_ptr())
    tr.transp_r2lm_dev_n(2, s_R.data_ptr(), s_LM.data_ptr())
    tr.transp_r2lm_dev_n(5, field_R.data_ptr(), field_LM.data_ptr())
    ext.synchronize()
    dflow_R, ds_R, db_R = calloc(3, nr_loc, lm_max), calloc(2, nr_loc, lm_max), calloc(3, nr_loc, lm_max)
    dflow_LM, ds_LM, db_LM = calloc(3, n_r_max, nlm_loc), calloc(2, n_r_max, nlm_loc), calloc(3, n_r_max, nlm_loc)
    dtr = torch.zeros(nr_loc, dtype=torch.float64, device=dev)
    dth = torch.zeros(nr_loc, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    p = make_params(physics, n_r_max)
    rad = make_radial(n_r_max, gs["l_max"], nRstart=tr.nRstart, nRstop=tr.nRstop)
    chunk = args.level_chunk
    if chunk == 0 and gs["l_max"] >= 1000:
        chunk = 16  # the ncu-profiled shape (also what the library's auto rule picks at this truncation)
    rl = RadialLoop(sht, p, rad, level_chunk=chunk)

    fin = {"w": flow_R[0], "dw": flow_R[1], "ddw": flow_R[2], "z": flow_R[3], "dz": flow_R[4], "s": s_R[0],
           "b": field_R[0], "db": field_R[1], "ddb": field_R[2], "aj": field_R[3], "dj": field_R[4]}
    fout = {"dwdt": dflow_R[0], "dzdt": dflow_R[1], "dpdt": dflow_R[2], "dsdt": ds_R[0], "dVSrLM": ds_R[1],
            "dbdt": db_R[0], "djdt": db_R[1], "dVxBhLM": db_R[2]}
    fin_p = {k: v.data_ptr() for k, v in fin.items()}
    fout_p = {k: v.data_ptr() for k, v in fout.items()}

    stage_acc = {}

    tev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tacc = {"transp_lm2r": 0.0, "transp_r2lm": 0.0}

    overlap = args.overlap == "on" or (args.overlap == "auto" and world > 1)

    def step_overlapped():
        # one call: the all-to-alls of level chunk c+1 (in) and c-1 (out) run on a second stream under the compute of chunk c
        rl.run_lm_dev(tr, flow_LM.data_ptr(), s_LM.data_ptr(), field_LM.data_ptr(), dflow_LM.data_ptr(), ds_LM.data_ptr(),
                      db_LM.data_ptr(), dtr.data_ptr(), dth.data_ptr())
        for k, v in rl.last_timing().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        pending.append(None)

    def step():
        if overlap:
            return step_overlapped()
        tev[0].record(ext)
        tr.transp_lm2r_dev_n(5, flow_LM.data_ptr(), flow_R.data_ptr())
        tr.transp_lm2r_dev_n(2, s_LM.data_ptr(), s_R.data_ptr())
        tr.transp_lm2r_dev_n(5, field_LM.data_ptr(), field_R.data_ptr())
        tev[1].record(ext)
        rl.radialLoop_dev(fin_p, fout_p, dtr.data_ptr(), dth.data_ptr())
        tev[2].record(ext)
        tr.transp_r2lm_dev_n(3, dflow_R.data_ptr(), dflow_LM.data_ptr())
        tr.transp_r2lm_dev_n(2, ds_R.data_ptr(), ds_LM.data_ptr())
        tr.transp_r2lm_dev_n(3, db_R.data_ptr(), db_LM.data_ptr())
        tev[3].record(ext)
        for k, v in rl.last_timing().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        # (the radial loop synchronises per chunk, so the lm2r events of this step have completed)
        tev[1].synchronize()
        tacc["transp_lm2r"] += tev[0].elapsed_time(tev[1])
        pending.append(None)

    pending = []

    def barrier():
        ext.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    stage_acc.clear()
    pending.clear()
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
        # 'total' spans the chunk loop including its waits for inbound chunks; what is left is the exposed tail of the last r2lm
        stages["transp_exposed_tail"] = ms - stages["total"]
    else:
        stages["transp_lm2r"] = tacc["transp_lm2r"] / max(len(pending), 1)
        stages["transp_r2lm_plus_wait"] = ms - stages["transp_lm2r"] - stages["total"]
    leg_ms = stages["legendre_syn"] + stages["legendre_an"]
    leg_tflops = rl.legendre_flops() / (leg_ms * 1e-3) * 1e-12 if leg_ms > 0 else 0.0
    checksum = float(torch.view_as_real(dflow_LM).abs().sum().item())

    # ---- end to end through the host-buffer C-ABI call (what the Fortran rIter_cuda_t binds) ------------------
    e2e = None
    if not args.no_e2e:
        del flow_LM, s_LM, field_LM, dflow_LM, ds_LM, db_LM
        host_in = {k: torch.empty(nr_loc, lm_max, dtype=torch.complex128).pin_memory() for k in fin}
        for k in fin:
            host_in[k].copy_(fin[k])
        del fin, fin_p, flow_R, s_R, field_R, fout, fout_p, dflow_R, ds_R, db_R
        torch.cuda.empty_cache()
        host_out = {k: torch.empty(nr_loc, lm_max, dtype=torch.complex128).pin_memory() for k in OUT_NAMES}
        # the host-buffer path pipelines H2D / compute / D2H over level chunks: give it at least four chunks
        e2e_chunk = chunk if (chunk and nr_loc // chunk >= 4) else max(4, nr_loc // 4)
        if e2e_chunk != chunk:
            rl.finalize()
            rl = RadialLoop(sht, p, rad, level_chunk=e2e_chunk)
        np_in = {k: v.numpy() for k, v in host_in.items()}
        np_out = {k: v.numpy() for k, v in host_out.items()}
        np_out["dtrkc"] = np.zeros(nr_loc)
        np_out["dthkc"] = np.zeros(nr_loc)
        h2d = sum(v.nbytes for v in np_in.values())
        d2h = sum(np_out[k].nbytes for k in ["dwdt", "dzdt", "dpdt", "dsdt", "dVSrLM", "dbdt", "djdt", "dVxBhLM", "dtrkc", "dthkc"])
        for _ in range(max(1, min(args.warmup, 2))):
            rl.radialLoop(np_in, out=np_out)
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(ext):
            e0.record(ext)
            for _ in range(args.steps):
                rl.radialLoop(np_in, out=np_out)
            e1.record(ext)
        barrier()
        wall = (time.perf_counter() - t0) / args.steps * 1e3
        ems = torch.tensor([max(e0.elapsed_time(e1) / args.steps, wall)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": total_flops / (float(ems.item()) * 1e-3) * 1e-9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(ems.item()), "level_chunk": e2e_chunk,
               "path": "magic_rloop_run (host R-distributed containers in, explicit terms out; transposes stay on the host side)"}

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
            "data": "synthetic: torch.randn seeded 20261017+1000*config+rank, Re,Im~N(0,1)/(l+1), Im(m=0)=0; random-init, no checkpoint",
            "config": config_dict(args, gs, rl and (chunk or "auto")),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "legendre_gemm_kernel (FP64 DMMA.8x8x4)", "achieved": leg_tflops,
                         "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": leg_tflops / FP64_PEAK_TFLOPS,
                         # dram__bytes_read+write per launch for a 16-level chunk at l_max=1023, mean of the synthesis
                         # (8.60 + 5.58 GB) and the analysis launch (13.37 + 1.34 GB) of
                         # profiles/r01/ncu_all_l1023_details_session2.csv; null for shapes that were not captured
                         "traffic": 14.45e9 if (gs["l_max"] == 1023 and chunk in (0, 16)) else None,
                         "peak_source": "measured DMMA m8n8k4 loop, tools/fp64_peak.cu -> profiles/fp64_peak_r01.json "
                                        "(MEASURED_PEAKS.json carries no FP64 figure)",
                         "share_of_step": leg_ms / ms_step},
            "cpu_baseline": cpu,
            "stages_ms": stages, "checksum": checksum, "sht_tflops_overall": value * 1e-3,
        }
        print(json.dumps(line), flush=True)
    rl.finalize()
    tr.destroy_comm()
    sht.finalize_sht()
    if world > 1:
        dist.destroy_process_group()


def main():
    # rank 0 prints ONE JSON line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION/INFO in some images) off it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO") and not os.environ.get("BENCH_KEEP_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = "WARN"
    args = parse()
    from magic_b200.workload import config_sizes
    gs = config_sizes(args.workload)
    if args.impl == "reference":
        run_reference(args, gs)
    else:
        run_magic(args, gs)


if __name__ == "__main__":
    main()
