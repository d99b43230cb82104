#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: G pair-interactions/s and % FP32
peak at N=1M; 1/2/4/8 B200 vs the reference's CPU implementation on the host cores).

    python bench.py --gpus 1 --steps K --warmup W            # N = 1,048,576 Plummer, fp32 (configs[2])
    torchrun --nproc-per-node G ... bench.py --gpus G ...    # N = 4,194,304 Plummer, strong scaling (configs[3])
    python bench.py --impl reference ...                     # the reference's own CPU code, all host threads

A "step" is one pass of the hot path: all-pairs force accumulation on the current positions plus
the kick-drift update, for every body.  `value` = N^2 * K / (device time, max over ranks), inputs
resident in HBM.  `e2e` = the same through the C-ABI with HOST buffers: every step uploads the
Body array from pinned memory, steps, and downloads it again, all inside the timed region.
Prints exactly one JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pair_interactions_per_s"
UNIT = "G pair-interactions/s"
FLOP_PER_INTERACTION = 20.0          # the convention both sides use
EPS = 0.01                           # Plummer softening of the synthetic configs
DT = 1.0e-3
SEED = 20260101


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--bodies", dest="n", type=int, default=0,
                    help="override body count (default 1M at 1 GPU, 4M at >1)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--splits", type=int, default=0)
    ap.add_argument("--exchange", choices=["auto", "nccl", "peer"], default="auto",
                    help="multi-GPU position exchange: auto = fused integrate-and-push over CUDA IPC peer mappings when "
                         "every rank can map its peers (else ncclAllGather); nccl / peer force one")
    return ap.parse_args()


def workload(args):
    n = args.n if args.n > 0 else (1048576 if args.gpus <= 1 else 4194304)
    name = (f"Plummer sphere N={n:,}, all-pairs force + kick-drift, fp32, eps={EPS}, dt={DT} "
            f"(BASELINE configs[{2 if n == 1048576 else 3 if n == 4194304 else '-'}])")
    return n, name


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.25):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in timed region"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def pinned_bodies(n):
    import torch
    from nbodysim_b200 import BODY_DTYPE

    t = torch.empty(n * 64, dtype=torch.uint8, pin_memory=True)
    return t, t.numpy().view(BODY_DTYPE)


def ncu_traffic_bytes():
    """DRAM bytes per force launch from the committed ncu capture (None if absent)."""
    try:
        tot = 0.0
        for line in open(os.path.join(ROOT, "profiles", "r1_force_uniform_ncu.txt")):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
        return tot or None
    except Exception:
        return None


def fp32_peak_tflops(info, sm_max_mhz):
    """B200 non-tensor FP32 peak: SMs x 128 FP32 lanes x 2 flop (FMA) x SM clock."""
    mhz = sm_max_mhz if sm_max_mhz else info["sm_clock_khz"] / 1e3
    return info["sm_count"] * 128 * 2 * mhz * 1e6 / 1e12, mhz


# ---------------------------------------------------------------------------------------------
def cpu_reference_rate(n, seconds, bodies=None):
    """The reference's own direct-sum code (Quadtree::acc leaf loop driven like Simulation::attract,
    oracle/_ref fast build = the reference's flags) on ALL host threads over a bounded sample of
    M targets x N sources of the same workload.  Falls back to the plain-C oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from nbodysim_b200 import ic

    if bodies is None:
        bodies = ic.plummer(n, seed=SEED, dims=2)   # the reference is 2-D: planar variant of the workload
    R = O.reference("fast")
    kind = "reference" if R is not None else "port"
    cores = os.cpu_count() or 1

    def run(m):
        out = np.zeros((m, 2), dtype=np.float32)
        t = time.perf_counter()
        if R is not None:
            R.ref_allpairs_acc(bodies.ctypes.data, n, EPS, 0, m, out.ctypes.data, 0)
        else:
            O.oracle().orc_allpairs_acc(bodies.ctypes.data, n, EPS, 2, 0, m, out.ctypes.data)
        return time.perf_counter() - t

    m0 = max(cores, 64)
    t0 = run(m0)                                      # calibration (also warms caches / threads)
    m1 = int(min(max(m0, m0 * 1.5 / max(t0, 1e-6)), 65536, n))   # second stage: ~1.5 s, thread start-up amortised
    m1 = max(cores, (m1 // cores) * cores)
    t1 = run(m1)
    m = int(min(max(m0, m1 * seconds / max(t1, 1e-6)), 65536, n))
    m = max(cores, (m // cores) * cores)
    return {"m": m, "n": n, "run": run, "kind": kind, "cores": cores,
            "threads_used": cores if R is not None else min(cores, int(os.environ.get("OMP_NUM_THREADS", cores)))}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, wname = workload(args)
    cal = cpu_reference_rate(n, seconds=max(1.0, min(6.0, 120.0 / max(1, args.steps + args.warmup))))
    m, run = cal["m"], cal["run"]
    for _ in range(args.warmup):
        run(m)
    t = [run(m) for _ in range(args.steps)]
    total = sum(t)
    value = m * n * args.steps / total / 1e9
    sample = f"{m} targets x {n} sources per step (of {n} x {n}); reference direct-sum leaf loop, {cal['cores']} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wname, "n": n, "sample": sample, "host": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cal["cores"], "kind": cal["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def native_arm(args):
    import torch
    from nbodysim_b200 import Simulation, capi, ic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    n, wname = workload(args)
    pin_t, host = pinned_bodies(n)
    host[:] = ic.plummer(n, seed=SEED, dims=3)

    kw = dict(dt=DT, eps=EPS, dims=3, device_ids=[local], j_splits=args.splits)
    if world > 1:
        idbuf = torch.zeros(capi.NBODY_NCCL_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            raw = (ctypes.c_uint8 * capi.NBODY_NCCL_ID_BYTES)()
            rc = capi.gpu_lib().nbody_gpu_nccl_unique_id(raw)
            if rc != 0:
                raise SystemExit(f"nbody_gpu_nccl_unique_id failed: {rc}")
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        idbuf = idbuf.to(dev)
        dist.broadcast(idbuf, 0)
        kw.update(world=world, rank=rank, nccl_id=bytes(idbuf.cpu().tolist()),
                  exchange={"auto": 0, "nccl": 1, "peer": 2}[args.exchange])
    stream = torch.cuda.current_stream()
    if world == 1:
        kw["stream"] = stream.cuda_stream          # launch on torch's current stream: torch events see the kernels
    sim = Simulation(host, **kw)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # 2x the 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sim.sync()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput -------------------------------------------------
    def measure():
        for _ in range(max(3, args.warmup)):
            sim.step(1)
        barrier()
        i0 = sim.info()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        step_ms, force_ms, integ_ms = [], [], []
        barrier()
        t_wall0 = time.time()
        for _ in range(args.steps):
            flush.fill_(1)                       # evict the L2-resident source array between timed steps
            sim.sync(); torch.cuda.synchronize()
            sim.profile_next_step(True)
            t0 = time.perf_counter()
            sim.step(1)                          # profiled step: CUDA events on the launch stream inside the library
            sim.sync()
            t1 = time.perf_counter()
            inf = sim.info()
            force_ms.append(inf["last_force_ms"]); integ_ms.append(inf["last_integ_ms"])
            # device time of the whole step = force + integrator kernels (+ allgather wait when distributed, which
            # the host-side wall clock around a synchronised step bounds from above)
            step_ms.append(inf["last_force_ms"] + inf["last_integ_ms"] if world == 1 else 1e3 * (t1 - t0))
        barrier()
        t_wall1 = time.time()
        i1 = sim.info()
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
        total_ms = sum(step_ms)
        if dist is not None:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, force_ms, integ_ms, i0, i1, clocks

    def throttled(clocks):
        if not clocks or not clocks.get("sm_mhz"):
            return False
        bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clocks.get("reasons", []))
        stuck = clocks["sm_mhz"] < 0.75 * clocks["sm_max_mhz"] and not clocks.get("reasons")
        return bool(bad) or stuck

    total_ms, force_ms, integ_ms, i0, i1, clocks = measure()
    remeasured = False
    flag = torch.tensor([1.0 if (rank == 0 and throttled(clocks)) else 0.0], device=dev)
    if dist is not None:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if flag.item() > 0:                              # thermal / hw slowdown or a leftover clock lock: take it once more
        remeasured = True
        time.sleep(2.0)
        total_ms, force_ms, integ_ms, i0, i1, clocks = measure()
    value = n * n * args.steps / (total_ms * 1e-3) / 1e9
    launches = i1["kernel_launches"] - i0["kernel_launches"]

    # ---------------- end to end through the C-ABI with host buffers ------------------------------
    e2e = None
    if not args.no_e2e:
        ke = args.steps
        sim.upload(host); sim.step(1); sim.download(out=host)     # warm the path
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            sim.upload(host)                  # H2D of the 64-byte Body array (pinned)
            sim.step(1)
            sim.download(out=host)            # D2H of pos/vel/acc of this rank's shard
        barrier()
        te = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        shard = i1["shard_count"] if world > 1 else n
        e2e = {"value": n * n * ke / te / 1e9, "unit": UNIT, "h2d_bytes_per_step": n * 64 * world,
               "d2h_bytes_per_step": min(shard, n) * 64 * world, "steps": ke, "ms_per_step": 1e3 * te / ke}

    # ---------------- the general-mass form of the same kernel, for transparency -------------------
    # The synthetic Plummer workload has equal masses, so the library runs the uniform-mass form (11 fp32
    # lane-ops per interaction).  Time the general form (12 lane-ops: the per-source mass multiply stays in
    # the loop) on the same bodies as well, 3 profiled steps after 2 warm-up steps.
    general = None
    if world == 1 and i1["uniform_mass"]:
        with Simulation(host, **dict(kw, force_variant=0)) as sg:
            sg.step(2)
            gf = []
            for _ in range(3):
                flush.fill_(1); sg.sync(); torch.cuda.synchronize()
                sg.profile_next_step(True)
                sg.step(1)
                gi = sg.info()
                gf.append(gi["last_force_ms"] + gi["last_integ_ms"])
            gms = statistics.mean(gf)
            general = {"value": n * n / (gms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": gms, "steps": 3,
                       "note": "force_variant=0: general-mass form (12 lane-ops/interaction), same bodies"}

    if rank != 0:
        sim.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (force) ------------------------------------
    peak_tf, peak_mhz = fp32_peak_tflops(i1, clocks.get("sm_max_mhz") if clocks else None)
    f_ms = statistics.mean(force_ms)
    per_gpu_inter = n * n / world
    achieved_tf = per_gpu_inter * FLOP_PER_INTERACTION / (f_ms * 1e-3) / 1e12
    roof = {
        "bound": "fp32", "kernel": "force_f32_fast_kernel", "achieved": achieved_tf, "peak": peak_tf,
        "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "peak_source": (f"computed: {i1['sm_count']} SMs x 128 FP32 lanes x 2 flop x {peak_mhz:.0f} MHz "
                        "(clocks.max.sm); MEASURED_PEAKS.json carries only HBM and bf16-tensor peaks and this "
                        "kernel uses neither"),
        "flop_per_interaction": FLOP_PER_INTERACTION, "interactions_per_launch_set": per_gpu_inter,
        "kernel_ms": f_ms, "traffic": None,
    }
    tr = ncu_traffic_bytes()
    if tr is not None and n == 1048576 and world == 1:
        roof["traffic"] = tr
        roof["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this N from the committed "
                                  "ncu --set full capture profiles/r1_force_uniform_ncu.txt")
    # bytes the kernel must move per launch: every source once (16 B) + one 12-byte partial per target and split
    roof["algorithmic_bytes"] = n * 16 + (n / world) * 12 * max(1, i1["j_splits"])
    if clocks and clocks.get("sm_mhz"):
        roof["frac_at_observed_clock"] = achieved_tf / (peak_tf * clocks["sm_mhz"] / peak_mhz)
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm = mp["hbm_gbs"]; src = "measured"
    except Exception:
        hbm, src = 6650.0, "fallback"
    i_ms = statistics.mean(integ_ms)
    nslots = max(1, i1["j_splits"])
    integ_bytes = (n / world) * (16 + 12 + 12 * nslots + 16 + 12 + 12)
    roof_integ = {"bound": "hbm", "kernel": "integrate_f32_kernel", "achieved": integ_bytes / (i_ms * 1e-3) / 1e9 if i_ms > 0 else None,
                  "peak": hbm, "unit": "GB/s", "frac": (integ_bytes / (i_ms * 1e-3) / 1e9 / hbm) if i_ms > 0 else None,
                  "peak_source": src, "bytes_per_body": integ_bytes / (n / world), "kernel_ms": i_ms}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cal = cpu_reference_rate(n, args.cpu_seconds)
        t = cal["run"](cal["m"])
        cpu = {"value": cal["m"] * n / t / 1e9, "unit": UNIT, "cores": cal["cores"], "kind": cal["kind"],
               "sample": f"{cal['m']} targets x {n} sources (of {n} x {n}), {t:.1f} s; reference direct-sum leaf loop "
                         f"(Quadtree.hpp:133-144) threaded like Simulation::attract, -O3 -ffast-math -march=x86-64-v3, "
                         f"planar variant of the same Plummer workload"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wname, "n": n, "dims": 3, "eps": EPS, "dt": DT, "ic_seed": SEED,
                   "rsqrt": "fast (MUFU.RSQ)", "parallelism": (f"targets sharded over {world} GPU(s); new positions reach the other ranks by " +
                                   ("peer stores from the integrator kernel over NVLink (CUDA IPC mappings, completion "
                                    "flags in peer memory; no collective call)" if i1["p2p_exchange"] == 2 else
                                    "ncclAllGather on a communication stream") if world > 1 else "one GPU"),
                   "exchange": {0: "nccl_allgather", 1: "peer_stores_one_process", 2: "peer_stores_ipc"}[i1["p2p_exchange"]] if world > 1 else None,
                   "l2": "flushed (256 MiB write) before every timed step", "j_splits": i1["j_splits"],
                   "force_ctas": i1["force_ctas"], "ctas_per_sm": i1["ctas_per_sm"], "fused_integrator": bool(i1["fused"]),
                   "mass_form": ("uniform-mass (equal masses detected: 11 fp32 lane-ops per interaction)" if i1["uniform_mass"]
                                 else "general masses (12 fp32 lane-ops per interaction)")},
        "tflops_20flop": value * FLOP_PER_INTERACTION / 1e3,
        "clocks": dict(clocks or {}, remeasured=remeasured), "e2e": e2e, "gpu_launches": launches, "roofline": roof, "roofline_integrator": roof_integ,
        "cpu_baseline": cpu, "general_mass_form": general,
    }
    if general is not None:
        general["frac_fp32_peak"] = general["value"] * 1e9 * FLOP_PER_INTERACTION / 1e12 / peak_tf
    print(json.dumps(line), flush=True)
    sim.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        native_arm(a)
