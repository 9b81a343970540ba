#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

metric : "1q-gate HBM GB/s at 30q (frac of peak); QFT-n sec/gate at 1/2/4/8 GPU"
N = 1  : configs[1] -- H / RX(1.0) / RZ(1.0) applied to every target qubit of a 30-qubit register, one
         unfused HBM pass per gate (90 passes per step).  value = gates * 32 * 2^n bytes / time  [GB/s].
         The same line carries `qft` (QFT-30 through QuantumCircuit::execute, fused and unfused, sec/gate),
         `roofline` (dominant kernel vs the measured HBM peak), `cpu_baseline` (the oracle's OpenMP port
         of the reference's rayon path on this box's host cores) and `e2e` (host buffers -> C ABI -> host).
N > 1  : weak scaling -- 30 local qubits per GPU, n = 30 + log2(N) total; the same sweep now includes the
         log2(N) global qubits, whose gates are a fused half-shard exchange over NVLink (dist.cu).
--impl reference : the reference's CPU path (oracle port; the Rust crate cannot be built in this image)
         on all host threads, same metric, bounded sample per step.

Timing: CUDA events on the engine's stream (spz_timer_start/stop), barrier + synchronize on both sides,
max over ranks.  Inputs (34 GB) are far larger than L2 (126 MB), so no explicit flush is needed.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SWEEP_GATES = (("H", ()), ("RX", (1.0,)), ("RZ", (1.0,)))  # benches/benchmark.rs:89,95 angles


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception as e:  # nvidia-smi missing
            log("clock sampler unavailable:", e)
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's rayon path on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_sweep_sample(n: int, threads: int, targets, reps: int = 1):
    """Times H/RX/RZ on `targets` of an n-qubit state with the oracle (OpenMP mirrors rayon's chunking).
    Returns (GB/s, seconds, gates)."""
    import numpy as np
    import oracle as orc
    orc.set_threads(threads)
    s = orc.State(n)
    amp = 1.0 / math.sqrt(1 << n)
    s.reals.fill(amp * math.cos(0.3))
    s.imags.fill(amp * math.sin(0.3))
    kinds = {"H": orc.H, "RX": orc.RX, "RZ": orc.RZ}
    orc.apply(orc.H, s, 0)  # touch pages / warm the thread pool
    t0 = time.perf_counter()
    gates = 0
    for _ in range(reps):
        for name, p in SWEEP_GATES:
            for t in targets:
                orc.apply(kinds[name], s, t, p)
                gates += 1
    dt = time.perf_counter() - t0
    return gates * 32.0 * (1 << n) / dt / 1e9, dt, gates


def pick_cpu_n(want: int) -> int:
    try:
        avail = 0
        for l in open("/proc/meminfo"):
            if l.startswith("MemAvailable"):
                avail = int(l.split()[1]) * 1024
        n = want
        while n > 20 and (16 << n) * 1.3 > avail:
            n -= 1
        return n
    except Exception:
        return min(want, 28)


def cpu_targets(n: int):
    return sorted({0, n // 2, n - 1})


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    import oracle as orc
    threads = orc.max_threads()
    n_total = 30 + int(math.log2(max(args.gpus, 1)))
    n = pick_cpu_n(min(n_total, args.cpu_qubits or 30))
    targets = cpu_targets(n)
    sample = (f"oracle port (C + OpenMP mirroring rayon chunking, gates.rs:361-372) of H/RX(1.0)/RZ(1.0) on targets "
              f"{targets} of a {n}-qubit state = {3 * len(targets)} unfused passes per step; full workload is all "
              f"{n_total} targets at {n_total} qubits")
    for _ in range(args.warmup):
        cpu_sweep_sample(n, threads, targets[:1])
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt, _g = cpu_sweep_sample(n, threads, targets)
        vals.append(v); secs.append(dt)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "1q_gate_effective_hbm_GBps", "value": value, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"sweep_1q_H_RX_RZ_all_targets_n{n_total}", "qubits": n_total, "sample_qubits": n},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is Rust (nightly) and cannot be compiled in this image; this is the oracle's C/OpenMP "
                "restatement of its rayon path (parallel over 2^(n-1-t) chunks, so target n-1 runs on one thread)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override local qubits per GPU (default 30)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip per-target table and QFT")
    ap.add_argument("--cpu-qubits", type=int, default=0, help="cap the qubit count of the CPU (reference / cpu_baseline) sample")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        log("warmup < 3 violates the timing rules; using 3")
        args.warmup = 3

    import numpy as np
    import spinoza_b200 as sb
    from spinoza_b200 import Gate, QuantumCircuit

    if sb.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device visible and there is no CPU fallback")

    dist = None
    if world > 1:
        from spinoza_b200 import distributed as sbd
        dist = sbd.init_from_env()
    n_local = args.qubits or 30
    g = int(math.log2(world))
    n = n_local + g
    peak, peak_src = measured_peaks()

    # device state
    free_b, total_b = sb.mem_info(local_rank)
    while 16 * (1 << n_local) > 0.9 * free_b:
        n_local -= 1
        n = n_local + g
    if dist is not None:
        state = sbd.DistState(n, dist)
    else:
        state = sb.State(n, device=local_rank)
    state.init_random(42)
    gates = [getattr(Gate, name) if not p else getattr(Gate, name)(*p) for name, p in SWEEP_GATES]
    targets = list(range(n))
    gates_per_step = len(gates) * len(targets)
    bytes_per_gate_total = 32.0 * (1 << n)          # all GPUs
    bytes_per_gate_gpu = 32.0 * (1 << n_local)      # per GPU

    def barrier():
        if dist is not None:
            dist.barrier()

    def apply(gate, t):
        sb.apply(gate, state, t)  # same call for a single-GPU State and for one shard of a DistState

    def step():
        for gate in gates:
            for t in targets:
                apply(gate, t)

    for _ in range(args.warmup):
        step()
    state.sync(); barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = sb.launch_count()
    state.sync(); barrier()
    state.timer_start()
    for _ in range(args.steps):
        step()
    ms = state.timer_stop()
    state.sync(); barrier()
    launches = sb.launch_count() - l0
    clocks = sampler.stop()
    nvlink = None
    if dist is not None:
        ms = dist.max_float(ms)
        st_ = state.stats()
        nvlink = {"exchanges_total": st_["exchanges"], "bytes_sent_per_gpu": st_["bytes_sent"],
                  "exchange_kernel_ms": st_["exchange_ms"], "GBps_per_direction_per_gpu": st_["nvlink_GBps_per_direction"],
                  "peak_GBps_per_direction": 770.0, "peak_source": "B200_PROFILING.md measured peer copy",
                  "frac": (st_["nvlink_GBps_per_direction"] or 0.0) / 770.0,
                  "note": "covers warm-up + timed steps; each exchange moves half a shard out and half a shard in"}
    ms_per_step = ms / args.steps
    value = gates_per_step * bytes_per_gate_total / (ms_per_step * 1e-3) / 1e9
    per_gate_ms = ms_per_step / gates_per_step
    achieved_gpu = bytes_per_gate_gpu / (per_gate_ms * 1e-3) / 1e9

    line = {
        "metric": "1q_gate_effective_hbm_GBps", "value": value, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"sweep_1q_H_RX_RZ_all_targets_n{n}", "qubits": n, "local_qubits": n_local,
                   "gates_per_step": gates_per_step, "state": "seeded random normalised (utils.rs:168-201 recipe, seed 42)",
                   "l2": "inputs (34 GB per GPU) are larger than L2; no flush needed", "fusion": "off (one HBM pass per gate)"},
        "roofline": {"bound": "hbm", "achieved": achieved_gpu, "peak": peak, "unit": "GB/s", "frac": achieved_gpu / peak,
                     "frac_of_nominal_8000": achieved_gpu / 8000.0, "peak_source": peak_src,
                     "kernel": "k_pair_vec / k_pair_low (kernels_direct.cuh)",
                     "algorithmic_bytes_per_launch": bytes_per_gate_gpu, "avg_launch_ms": per_gate_ms,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one k_pair_vec launch at n = 30, from the committed
                     # ncu --set full capture (profiles/round1_k_pair_vec_ncu_full_raw.csv); only valid for 30 local qubits
                     "traffic": 34.30e9 if n_local == 30 else None,
                     "traffic_source": "profiles/round1_k_pair_vec_ncu_full_raw.csv (17.18 GB read + 17.12 GB written)"},
        "clocks": clocks, "gpu_launches": int(launches),
    }
    if nvlink is not None:
        line["nvlink"] = nvlink

    # ---- per-(gate, target) table, single GPU only: 1 warm-up + 5 timed reps each ----
    if rank == 0 and dist is None and not args.no_extras:
        table = {}
        for (name, _p), gate in zip(SWEEP_GATES, gates):
            row = []
            for t in targets:
                apply(gate, t)
                state.timer_start()
                for _ in range(5):
                    apply(gate, t)
                row.append(bytes_per_gate_gpu * 5 / (state.timer_stop() * 1e-3) / 1e9)
            table[name] = [round(x, 1) for x in row]
        flat = [x for r in table.values() for x in r]
        worst = min(((x, nm, t) for nm, r in table.items() for t, x in enumerate(r)))
        line["sweep"] = {"per_target_GBps": table, "min": min(flat), "median": statistics.median(flat), "max": max(flat),
                         "worst": f"{worst[1]}@t={worst[2]}", "frac_min_of_peak": min(flat) / peak,
                         "frac_median_of_peak": statistics.median(flat) / peak}

    # ---- QFT-n through QuantumCircuit::execute (fused and unfused), sec/gate ----
    def run_qft():
        qft = {}
        n_gates = n + n * (n - 1) // 2
        for label, fuse in (("fused", True), ("unfused", False)):
            state.set_basis(0x9E3779B97F4A7C15 % (1 << n))
            qc = QuantumCircuit.from_state(state, fuse=fuse)
            qc.qft()
            state.sync(); barrier()
            l1 = sb.launch_count()
            state.timer_start()
            qc.execute()
            t_ms = state.timer_stop()
            barrier()
            if dist is not None:
                t_ms = dist.max_float(t_ms)
            qft[label] = {"seconds": t_ms * 1e-3, "sec_per_gate": t_ms * 1e-3 / n_gates, "launches": int(sb.launch_count() - l1),
                          "effective_GBps_per_gpu": n_gates * bytes_per_gate_gpu / (t_ms * 1e-3) / 1e9}
        qft["gates"] = n_gates
        qft["qubits"] = n
        # which fused tile kernel ran: k_tile3 (csrc/kernels_tile3.cu, TMA; default) or k_tile (SPZ_TILE_V3=0)
        qft["tile_kernel"] = "k_tile" if os.environ.get("SPZ_TILE_V3", "1")[:1] == "0" else "k_tile3"
        # closed form check on a sample of amplitudes: QFT|x>[k] = 2^(-n/2) exp(2 pi i x rev(k) / 2^n)
        if dist is None:
            x = 0x9E3779B97F4A7C15 % (1 << n)
            re, im = state.download(0, 4096)
            k = np.arange(4096, dtype=np.uint64)
            rev = np.zeros_like(k)
            for b in range(n):
                rev |= ((k >> np.uint64(b)) & np.uint64(1)) << np.uint64(n - 1 - b)
            ph = ((np.uint64(x) * rev) % np.uint64(1 << n)).astype(np.float64) / float(1 << n)
            want = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ph)
            qft["max_abs_err_vs_closed_form_first_4096"] = float(np.max(np.abs((re + 1j * im) - want)))
        else:
            qft["norm2_after"] = sb.norm2(state)
            qft["exchanges_total_incl_sweep"] = state.stats()["exchanges"]
        return qft

    if not args.no_extras:
        try:
            line["qft"] = run_qft()
        except Exception as e:  # keep the headline line even if an extra fails
            line["qft"] = {"error": repr(e)}

    # ---- BASELINE config 3 through QuantumCircuit::execute (single GPU): random layered circuit, depth 20 ----
    # Timed twice: with the scheduler choosing each pass's tile qubits (default from 24 qubits up) and with the round-1 rule
    # (first ready ops claim the tile, SPZ_TILE_SELECT=0), so that one bench line shows what the shorter plan is worth.
    def run_config3():
        from spinoza_b200 import workloads
        out = {"qubits": n, "depth": 20}
        saved = os.environ.get("SPZ_TILE_SELECT")
        try:
            for label, select in (("chosen_tile", None), ("first_come_tile", "0")):
                if select is None:
                    os.environ.pop("SPZ_TILE_SELECT", None)
                else:
                    os.environ["SPZ_TILE_SELECT"] = select
                state.init_random(42)
                qc = QuantumCircuit.from_state(state, fuse=True)
                out["gates"] = workloads.random_layered_circuit(qc, depth=20, seed=42)
                passes = qc.plan()[1]
                state.sync()
                l1 = sb.launch_count()
                state.timer_start()
                qc.execute()
                t_ms = state.timer_stop()
                out[label] = {"seconds": t_ms * 1e-3, "passes": passes, "launches": int(sb.launch_count() - l1),
                              "sec_per_gate": t_ms * 1e-3 / out["gates"], "norm2_after": sb.norm2(state)}
        finally:
            if saved is None:
                os.environ.pop("SPZ_TILE_SELECT", None)
            else:
                os.environ["SPZ_TILE_SELECT"] = saved
        return out

    if rank == 0 and dist is None and not args.no_extras:
        try:
            line["config3"] = run_config3()
        except Exception as e:
            line["config3"] = {"error": repr(e)}

    # ---- end to end through the C ABI with HOST buffers: upload -> sweep -> download, every step ----
    # (sharded: every rank moves its own shard between its pinned host buffers and its GPU; max over ranks)
    def mem_available():
        try:
            for l in open("/proc/meminfo"):
                if l.startswith("MemAvailable"):
                    return int(l.split()[1]) * 1024
        except Exception:
            pass
        return 0

    if not args.no_e2e:
        try:
            nbytes = 8 << n_local
            if mem_available() < 2.5 * 2 * nbytes * world:
                raise MemoryError(f"host has {mem_available() >> 30} GiB available; e2e needs {(2 * nbytes * world) >> 30} GiB pinned")
            hre = sb.HostBuffer(1 << n_local)
            him = sb.HostBuffer(1 << n_local)
            state.init_random(42)
            state.download_into(hre, him)
            e2e_steps = max(1, min(args.steps, 2))
            state.upload_from(hre, him); step(); state.download_into(hre, him)  # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                state.upload_from(hre, him)
                step()
                state.download_into(hre, him)
            dt = (time.perf_counter() - t0) / e2e_steps
            if dist is not None:
                dt = dist.max_float(dt)
            line["e2e"] = {"value": gates_per_step * bytes_per_gate_total / dt / 1e9, "unit": "GB/s",
                           "h2d_bytes_per_step": 2 * nbytes * world, "d2h_bytes_per_step": 2 * nbytes * world,
                           "ms_per_step": dt * 1e3, "steps": e2e_steps,
                           "what": f"pinned host re/im -> spz_upload -> {gates_per_step} x spz_apply -> spz_download (whole state"
                                   + (", every rank its own shard" if dist is not None else "") + "), wall clock, max over ranks"}
            del hre, him
        except Exception as e:  # host cannot pin the buffers
            line["e2e"] = {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(e)}

    # ---- CPU baseline: oracle port on this box's host cores, bounded sample ----
    if rank == 0 and not args.no_cpu and world == 1:
        del state
        import oracle as orc
        threads = orc.max_threads()
        cn = pick_cpu_n(min(n, args.cpu_qubits or 30))
        tg = cpu_targets(cn)
        v, dt, gcount = cpu_sweep_sample(cn, threads, tg)
        v1, dt1, _ = cpu_sweep_sample(min(cn, 26), 1, cpu_targets(min(cn, 26)))
        line["cpu_baseline"] = {"value": v, "unit": "GB/s", "cores": threads, "kind": "port",
                                "sample": f"{gcount} unfused passes (H/RX/RZ on targets {tg}) of a {cn}-qubit state, {dt:.1f} s",
                                "single_thread_GBps": v1}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.shutdown()


if __name__ == "__main__":
    main()
