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


def cpu_targets(n: int, threads: int):
    """Weighted target sample for the CPU sweep.  The reference parallelises a pass over its 2^(n-1-t) chunks
    (gates.rs:351-372, 601-615), so target n-1 runs on ONE thread, n-2 on two, ...: the top log2(threads) targets are each
    timed themselves (weight 1); every lower target has at least `threads` chunks and is represented by four of them
    (0, n/3, 2n/3 and the highest fully parallel one), which share the remaining weight.  Returns [(target, weight)] with
    weights summing to n, so that sum(weight * seconds) estimates one full sweep over all n targets."""
    top = max(1, min(n - 1, int(math.ceil(math.log2(max(threads, 1))))))
    slow = list(range(n - top, n))
    n_fast = n - top
    reps = sorted({0, n_fast // 3, (2 * n_fast) // 3, n_fast - 1})
    return [(t, n_fast / len(reps)) for t in reps] + [(t, 1.0) for t in slow]


def cpu_sweep_weighted(n: int, threads: int, full: bool = False):
    """The 1-qubit sweep on the host cores with the oracle port.  Returns (GB/s over the whole sweep, measured seconds, passes
    timed, estimated seconds of the full sweep).  full=True times every target instead of the weighted sample."""
    import oracle as orc
    orc.set_threads(threads)
    s = orc.State(n)
    amp = 1.0 / math.sqrt(1 << n)
    s.reals.fill(amp * math.cos(0.3))
    s.imags.fill(amp * math.sin(0.3))
    kinds = {"H": orc.H, "RX": orc.RX, "RZ": orc.RZ}
    orc.apply(orc.H, s, 0)  # touch pages / warm the thread pool
    tw = [(t, 1.0) for t in range(n)] if full else cpu_targets(n, threads)
    measured = est = 0.0
    passes = 0
    for name, p in SWEEP_GATES:
        for t, w in tw:
            t0 = time.perf_counter()
            orc.apply(kinds[name], s, t, p)
            dt = time.perf_counter() - t0
            measured += dt
            est += w * dt
            passes += 1
    gates_full = len(SWEEP_GATES) * n
    return gates_full * 32.0 * (1 << n) / est / 1e9, measured, passes, est


def cpu_qft(n: int, threads: int):
    """BASELINE config 1: QFT-n through the oracle's execute (the reference's one-pass-per-gate loop, circuit.rs:553-599),
    seconds per gate."""
    import oracle as orc
    from spinoza_b200 import QuantumCircuit, QuantumRegister
    from spinoza_b200.circuit import Controls  # noqa: F401  (kept for symmetry with the parity tests)
    orc.set_threads(threads)
    qc = QuantumCircuit(QuantumRegister(n))
    qc.qft()
    ops = [orc.make_op(t.gate.kind, t.target, t.gate.params, ctrl_kind=t.controls.kind, ctrl_mask=t.controls.mask())
           for t in qc.transformations]
    s = orc.State(n)
    orc.execute(s, ops[:n])  # warm the thread pool and the pages
    s = orc.State(n)
    t0 = time.perf_counter()
    orc.execute(s, ops)
    dt = time.perf_counter() - t0
    return {"seconds": dt, "gates": len(ops), "sec_per_gate": dt / len(ops), "threads": threads}


CRITERION_N = 25  # benches/benchmark.rs:148


def _qcbm_ops(n, depth=9):
    """The op list of the reference's `qcbm` bench (workloads.qcbm) for the oracle's execute."""
    import oracle as orc
    from spinoza_b200 import QuantumCircuit, QuantumRegister, workloads
    qc = QuantumCircuit(QuantumRegister(n))
    workloads.qcbm(qc, depth=depth, seed=42)
    return [orc.make_op(t.gate.kind, t.target, t.gate.params, ctrl_kind=t.controls.kind, ctrl_mask=t.controls.mask())
            for t in qc.transformations]


def cpu_criterion(threads: int, n: int = CRITERION_N):
    """The reference's own criterion suite (benches/benchmark.rs:147-188, n = 25) on the oracle port: seconds per iteration of
    each bench function, built as the reference builds it (h / x / rx / p / z / u / value_encoding allocate their State inside
    the timed closure, `measure` generates its random state there).  One timed iteration after one warm-up."""
    import oracle as orc
    orc.set_threads(threads)
    PI = math.pi

    def fresh_loop(kind, p=()):
        def f():
            s = orc.State(n)
            for t in range(n):
                orc.apply(kind, s, t, p)
        return f

    kept = orc.State(n)
    pairs = [(i, (i + 1) % n) for i in range(n)]
    qcbm_ops = _qcbm_ops(n)
    qcbm_state = orc.State(n)

    def cx():
        for c, t in pairs:
            orc.c_apply(orc.X, kept, c, t)

    def rz():
        for t in range(n):
            orc.apply(orc.RZ, kept, t, (1.0,))

    def value_encoding():  # benches/benchmark.rs:63-76
        s = orc.State(n)
        for i in range(n):
            orc.apply(orc.H, s, i)
        for i in range(n):
            orc.apply(orc.P, s, i, (2.0 * PI / (2.0 ** (i + 1)) * 2.4,))
        orc.iqft(s, list(range(n - 1, -1, -1)))

    def measure():  # benches/benchmark.rs:142-145
        s = orc.gen_random_state(n, 42)
        orc.measure_qubit(s, 0, True, None, 0.5)

    fns = {"h": fresh_loop(orc.H), "x": fresh_loop(orc.X), "cx": cx, "rz": rz, "rx": fresh_loop(orc.RX, (1.0,)),
           "qcbm": lambda: orc.execute(qcbm_state, qcbm_ops), "p": fresh_loop(orc.P, (1.0,)), "z": fresh_loop(orc.Z),
           "u": fresh_loop(orc.U, (1.0, 2.0, 3.0)), "value_encoding": value_encoding, "measure": measure}
    out = {}
    fresh_loop(orc.H)()  # thread pool, page cache
    for name, f in fns.items():
        t0 = time.perf_counter()
        f()
        out[name] = time.perf_counter() - t0
    return out


def gpu_criterion(sb, device: int, n: int = CRITERION_N, reps: int = 5):
    """The same suite through this engine's mirror of the reference API on one GPU: median seconds per iteration over `reps`
    (CUDA events on the state's stream; `measure` by wall clock, it returns a value to the host).  The State is created once
    and reset (spz_reset_zero) where the reference allocates a fresh one per iteration.  qcbm / value_encoding's iqft go
    through QuantumCircuit::execute fused (the drop-in's default) -- qcbm_unfused is the one-pass-per-gate figure."""
    from spinoza_b200 import Gate, QuantumCircuit, workloads
    PI = math.pi
    st = sb.State(n, device=device)
    pairs = [(i, (i + 1) % n) for i in range(n)]

    def fresh_loop(gate):
        def f():
            st.reset()
            for t in range(n):
                sb.apply(gate, st, t)
        return f

    def cx():
        for c, t in pairs:
            sb.c_apply(Gate.X, st, c, t)

    def rz():
        for t in range(n):
            sb.apply(Gate.RZ(1.0), st, t)

    def qcbm(fuse):
        def f():
            qc = QuantumCircuit.from_state(st, fuse=fuse)
            workloads.qcbm(qc, depth=9, seed=42)
            return qc
        return f

    def value_encoding():
        st.reset()
        for i in range(n):
            sb.apply(Gate.H, st, i)
        for i in range(n):
            sb.apply(Gate.P(2.0 * PI / (2.0 ** (i + 1)) * 2.4), st, i)
        sb.iqft(st, list(range(n - 1, -1, -1)))

    out = {}

    def timed(name, body, prepare=None):
        ts = []
        for r in range(reps + 1):
            arg = prepare() if prepare else None
            st.sync()
            st.timer_start()
            body(arg) if prepare else body()
            ts.append(st.timer_stop() * 1e-3)
        out[name] = statistics.median(ts[1:])

    timed("h", fresh_loop(Gate.H))
    timed("x", fresh_loop(Gate.X))
    st.reset(); timed("cx", cx)
    st.reset(); timed("rz", rz)
    timed("rx", fresh_loop(Gate.RX(1.0)))
    st.reset(); timed("qcbm", lambda qc: qc.execute(), prepare=qcbm(True))
    st.reset(); timed("qcbm_unfused", lambda qc: qc.execute(), prepare=qcbm(False))
    timed("p", fresh_loop(Gate.P(1.0)))
    timed("z", fresh_loop(Gate.Z))
    timed("u", fresh_loop(Gate.U(1.0, 2.0, 3.0)))
    timed("value_encoding", value_encoding)
    ts = []
    for r in range(reps + 1):
        st.sync()
        t0 = time.perf_counter()
        st.init_random(42)
        sb.measure_qubit(st, 0, True, None)
        ts.append(time.perf_counter() - t0)
    out["measure"] = statistics.median(ts[1:])
    qc = qcbm(True)()
    out["_qcbm"] = {"gates": len(qc.transformations), "passes_fused": qc.plan()[1]}
    del st
    return out


def sweep_config(n: int, n_local: int):
    """The `config` object of the headline workload: identical in both arms."""
    return {"workload": f"sweep_1q_H_RX_RZ_all_targets_n{n}", "qubits": n, "local_qubits": n_local,
            "gates_per_step": len(SWEEP_GATES) * n, "state": "seeded random normalised (utils.rs:168-201 recipe, seed 42)",
            "l2": "inputs (34 GB per GPU) are larger than L2; no flush needed", "fusion": "off (one HBM pass per gate)"}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    import oracle as orc
    threads = orc.max_threads()
    n_local = args.qubits or 30
    n_total = n_local + int(math.log2(max(args.gpus, 1)))
    n = pick_cpu_n(min(n_total, args.cpu_qubits or 30))
    tw = cpu_targets(n, threads)
    sample = (f"oracle port (C + OpenMP mirroring rayon's chunking, gates.rs:351-372) of H/RX(1.0)/RZ(1.0) on a {n}-qubit state, "
              f"{3 * len(tw)} timed passes per step on the weighted target sample {[(t, round(w, 2)) for t, w in tw]} "
              f"(weights sum to {n}: the top targets, which the reference runs on 1, 2, 4, ... threads, are each timed themselves); "
              f"value = bytes of the full {n}-target sweep / weighted seconds.  GB/s does not depend on the register size, so the "
              f"{n}-qubit sample stands for the {n_total}-qubit workload")
    for _ in range(min(args.warmup, 1)):
        cpu_sweep_weighted(min(n, 26), threads)
    vals, secs = [], []
    for _ in range(args.steps):
        v, measured, _passes, est = cpu_sweep_weighted(n, threads)
        vals.append(v); secs.append(est)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "1q_gate_effective_hbm_GBps", "value": value, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": sweep_config(n_total, n_local),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample, "sample_qubits": n},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is Rust (nightly) and cannot be compiled in this image; this is the oracle's C/OpenMP "
                "restatement of its rayon path (parallel over 2^(n-1-t) chunks, so target n-1 runs on one thread); ms_per_step is "
                "the weighted estimate of one full sweep",
    }
    if not args.no_extras:
        try:  # BASELINE config 1: QFT-20 on the CPU at 1 thread and at all host threads
            line["qft20"] = {"one_thread": cpu_qft(20, 1), "all_threads": cpu_qft(20, threads)}
        except Exception as e:
            line["qft20"] = {"error": repr(e)}
        try:  # the reference's own criterion suite (benches/benchmark.rs:147-188) at its n = 25
            line["criterion"] = {"qubits": pick_cpu_n(CRITERION_N), "seconds": cpu_criterion(threads, pick_cpu_n(CRITERION_N)), "threads": threads}
        except Exception as e:
            line["criterion"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)



def qft_closed_form_err(np, state, n, x, env=None, sample=1 << 16):
    """max |amp - 2^(-n/2) exp(2 pi i x rev(k) / 2^n)| over the first `sample` amplitudes of this GPU's shard (physical
    index mapped back to the logical k through the qubit permutation); max over ranks when sharded."""
    cnt = min(sample, len(state))
    re, im = state.download(0, cnt)
    if env is not None:
        perm, n_local, rank = state.perm(), state.n_local, env.rank
    else:
        perm, n_local, rank = list(range(n)), n, 0
    phys = (np.uint64(rank) << np.uint64(n_local)) + np.arange(cnt, dtype=np.uint64)
    rev = np.zeros(cnt, dtype=np.uint64)
    for q in range(n):
        rev |= ((phys >> np.uint64(perm[q])) & np.uint64(1)) << np.uint64(n - 1 - q)
    ph = ((np.uint64(x) * rev) & np.uint64((1 << n) - 1)).astype(np.float64) / float(1 << n)  # wrap-around is exact mod 2^n
    want = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ph)
    err = float(np.max(np.abs((re + 1j * im) - want)))
    return env.max_float(err) if env is not None else err


def sharded_vs_oracle(np, sb, sbd, env, n=20):
    """Parity of the sharded engine on real NVLink: a 20-qubit register over all ranks runs the layered circuit followed by
    a QFT through the fused execute; every rank replays the same gate list with the CPU oracle and compares ITS shard
    (physical order mapped through the permutation).  Returns max |difference| over all ranks."""
    import oracle as orc
    from spinoza_b200 import QuantumCircuit, workloads
    init = orc.gen_random_state(n, 4242)
    s = sbd.DistState(n, env)
    nl = s.n_local
    lo, hi = env.rank << nl, (env.rank + 1) << nl
    s.upload(np.ascontiguousarray(init.reals[lo:hi]), np.ascontiguousarray(init.imags[lo:hi]))
    qc = QuantumCircuit.from_state(s, fuse=True)
    workloads.random_layered_circuit(qc, depth=10, seed=7)
    qc.qft()
    ops = [orc.make_op(t.gate.kind, t.target, t.gate.params, ctrl_kind=t.controls.kind, ctrl_mask=t.controls.mask())
           for t in qc.transformations]
    x0 = s.stats()["exchanges"]
    qc.execute()
    s.sync()
    orc.execute(init, ops)
    perm = s.perm()
    re, im = s.download()
    phys = (np.int64(env.rank) << np.int64(nl)) + np.arange(1 << nl, dtype=np.int64)
    logical = np.zeros_like(phys)
    for q in range(n):
        logical |= ((phys >> perm[q]) & 1) << q
    err = float(max(np.max(np.abs(re - init.reals[logical])), np.max(np.abs(im - init.imags[logical]))))
    out = {"qubits": n, "gates": len(ops), "exchanges": s.stats()["exchanges"] - x0, "max_abs_err": env.max_float(err)}
    del s
    return out


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def measured_traffic(n_local):
    p = ROOT / "profiles" / f"round2_traffic_n{n_local}.json"
    try:
        d = json.loads(p.read_text())
        return {"traffic": d["traffic_full_pass_mean"],
                "traffic_source": f"profiles/{p.name}: mean DRAM read + write of the H / RX launches, {d['source']}"}
    except Exception:
        return {"traffic": None, "traffic_source": f"no profiles/{p.name} (run tools/measure_traffic.py {n_local})"}


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
    ap.add_argument("--no-northstar", action="store_true", help="sharded runs: skip QFT at 2^33 amplitudes per GPU (BASELINE config 4)")
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
        "config": sweep_config(n, n_local),
        "roofline": {"bound": "hbm", "achieved": achieved_gpu, "peak": peak, "unit": "GB/s", "frac": achieved_gpu / peak,
                     "frac_of_nominal_8000": achieved_gpu / 8000.0, "peak_source": peak_src,
                     "kernel": "k_pair_vec / k_pair_low (kernels_direct.cuh)",
                     "algorithmic_bytes_per_launch": bytes_per_gate_gpu, "avg_launch_ms": per_gate_ms,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch of the sweep's kernels at this size, measured by
                     # tools/measure_traffic.py (an ncu pass, committed under profiles/); null when no measurement exists for it
                     **measured_traffic(n_local)},
        "clocks": clocks, "gpu_launches": int(launches),
    }
    if nvlink is not None:
        line["nvlink"] = nvlink

    # ---- per-(gate, target) table, single GPU only: 1 warm-up + 20 timed reps each ----
    if rank == 0 and dist is None and not args.no_extras:
        table = {}
        for (name, _p), gate in zip(SWEEP_GATES, gates):
            row = []
            for t in targets:
                apply(gate, t)
                state.timer_start()
                for _ in range(20):
                    apply(gate, t)
                row.append(bytes_per_gate_gpu * 20 / (state.timer_stop() * 1e-3) / 1e9)
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
            if fuse:  # one untimed run: the first fused execute pays one-time costs (kernel attributes, the TMA descriptor encoder)
                state.set_basis(1)
                warm = QuantumCircuit.from_state(state, fuse=True)
                warm.qft()
                warm.execute()
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
        # closed form check, QFT|x>[k] = 2^(-n/2) exp(2 pi i x rev(k) / 2^n), on the first 65 536 amplitudes of every shard
        x = 0x9E3779B97F4A7C15 % (1 << n)
        qft["max_abs_err_vs_closed_form"] = qft_closed_form_err(np, state, n, x, dist)
        qft["norm2_after"] = sb.norm2(state)
        if dist is not None:
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

    # ---- the reference's own criterion suite (benches/benchmark.rs:147-188) at its n = 25, single GPU ----
    if rank == 0 and dist is None and not args.no_extras:
        try:
            gc = gpu_criterion(sb, local_rank)
            meta = gc.pop("_qcbm")
            line["criterion"] = {"qubits": CRITERION_N, "gpu_seconds": gc, "qcbm": meta,
                                 "what": "benches/benchmark.rs:147-188 bench functions (one iteration each) through the mirrored API; "
                                         "median of 5 after one warm-up; the CPU port's figures are in cpu_baseline.criterion"}
        except Exception as e:
            line["criterion"] = {"error": repr(e)}

    # ---- reductions at the bench size (single GPU): <Z> of every qubit from one read pass vs one pass per qubit ----
    if rank == 0 and dist is None and not args.no_extras:
        try:
            state.init_random(42)
            sb.xyz_expectation_value("z", state, [0, 1])  # warm
            t0 = time.perf_counter(); zs = sb.xyz_expectation_value("z", state, list(range(n))); t_all = time.perf_counter() - t0
            t0 = time.perf_counter(); z1 = [sb.xyz_expectation_value("z", state, [t])[0] for t in range(n)]; t_each = time.perf_counter() - t0
            line["reductions"] = {"qubits": n, "z_all_qubits_one_pass_s": t_all, "z_one_pass_per_qubit_s": t_each,
                                  "effective_read_GBps_one_pass": 16.0 * (1 << n) / t_all / 1e9,
                                  "max_abs_diff": float(max(abs(a - b) for a, b in zip(zs, z1)))}
            # <X> of every qubit: tiles staged in shared memory, up to twelve targets per read pass, vs one pass per qubit
            sb.xyz_expectation_value("x", state, [0, 1, 2])  # warm (kernel attribute)
            t0 = time.perf_counter(); xs = sb.xyz_expectation_value("x", state, list(range(n))); t_xall = time.perf_counter() - t0
            t0 = time.perf_counter(); x1 = [sb.xyz_expectation_value("x", state, [t])[0] for t in range(n)]; t_xeach = time.perf_counter() - t0
            line["reductions"].update({"x_all_qubits_tiled_s": t_xall, "x_one_pass_per_qubit_s": t_xeach,
                                       "x_max_abs_diff": float(max(abs(a - b) for a, b in zip(xs, x1)))})
        except Exception as e:
            line["reductions"] = {"error": repr(e)}

    # ---- sharded runs: parity against the CPU oracle on real NVLink (20 qubits over all ranks) ----
    if dist is not None and not args.no_extras:
        try:
            line["parity"] = {"qft_closed_form_max_abs_err": line.get("qft", {}).get("max_abs_err_vs_closed_form"),
                              "sharded_vs_oracle": sharded_vs_oracle(np, sb, sbd, dist, 20),
                              "what": "QFT of a basis state against the closed form on 65 536 amplitudes of every shard; layered circuit + "
                                      "QFT at 20 qubits over all ranks against the CPU oracle, every amplitude; max over ranks"}
        except Exception as e:
            line["parity"] = {"error": repr(e)}

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
            e2e_steps = max(1, min(args.steps, 5))
            # spz_upload_async: the state arrives in four pieces on a copy stream and the gates on targets below the piece bits
            # follow piece by piece behind the bus; spz_download waits for everything (SPZ_BENCH_SYNC_UPLOAD=1: spz_upload)
            sync_upload = os.environ.get("SPZ_BENCH_SYNC_UPLOAD") == "1"
            upload = state.upload_from if sync_upload else state.upload_async
            upload(hre, him); step(); state.download_into(hre, him)  # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                upload(hre, him)
                step()
                state.download_into(hre, him)
            dt = (time.perf_counter() - t0) / e2e_steps
            if dist is not None:
                dt = dist.max_float(dt)
            line["e2e"] = {"value": gates_per_step * bytes_per_gate_total / dt / 1e9, "unit": "GB/s",
                           "h2d_bytes_per_step": 2 * nbytes * world, "d2h_bytes_per_step": 2 * nbytes * world,
                           "ms_per_step": dt * 1e3, "steps": e2e_steps,
                           "what": f"pinned host re/im -> {'spz_upload' if sync_upload else 'spz_upload_async'} -> {gates_per_step} x spz_apply -> spz_download (whole state"
                                   + (", every rank its own shard" if dist is not None else "") + "), wall clock, max over ranks"}
            del hre, him
        except Exception as e:  # host cannot pin the buffers
            line["e2e"] = {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(e)}

    # ---- BASELINE config 4 at its stated size: QFT-(33 + log2 N) with 2^33 amplitudes (137 GB) per GPU ----
    if dist is not None and not args.no_extras and not args.no_northstar:
        try:
            del state
            state = None
            nl = 33
            free_b, _tot = sb.mem_info(local_rank)
            fits = dist.max_float(0.0 if 16 * (1 << nl) * 1.01 < free_b else 1.0) == 0.0
            if not fits:
                line["qft_northstar"] = {"skipped": f"2^33 amplitudes need 137.4 GB per GPU; {free_b / 1e9:.1f} GB free"}
            else:
                nn = nl + g
                t0 = time.perf_counter()
                big = sbd.DistState(nn, dist)
                x = 0x9E3779B97F4A7C15 % (1 << nn)
                big.set_basis(x)
                big.sync(); barrier()
                alloc_s = time.perf_counter() - t0
                qc = QuantumCircuit.from_state(big, fuse=True)
                qc.qft()
                n_gates = len(qc.transformations)
                s0 = big.stats()
                l1 = sb.launch_count()
                big.sync(); barrier()
                big.timer_start()
                qc.execute()
                t_ms = dist.max_float(big.timer_stop())
                barrier()
                s1 = big.stats()
                xms = s1["exchange_ms"] - s0["exchange_ms"]
                sent = s1["bytes_sent"] - s0["bytes_sent"]
                line["qft_northstar"] = {
                    "workload": f"QFT-{nn} on {world} GPUs, 2^{nl} amplitudes (137.4 GB) per GPU, fused", "qubits": nn, "gates": n_gates,
                    "seconds": t_ms * 1e-3, "sec_per_gate": t_ms * 1e-3 / n_gates, "launches": int(sb.launch_count() - l1),
                    "exchanges": s1["exchanges"] - s0["exchanges"], "overlapped_exchanges": s1["overlapped"] - s0["overlapped"],
                    "exchange_kernel_ms": xms, "nvlink_GBps_per_direction_per_gpu": (sent / (xms * 1e-3) / 1e9) if xms > 0 else None,
                    "max_abs_err_vs_closed_form": qft_closed_form_err(np, big, nn, x, dist), "norm2_after": sb.norm2(big),
                    "alloc_seconds": alloc_s,
                    "placement": ("the register was still a basis state when execute started, so its qubit permutation was chosen from the "
                                  "op list (dist_place_basis): g exchanges instead of g + 1; SPZ_DIST_PLACE=0 restores the identity placement")
                                 if os.environ.get("SPZ_DIST_PLACE", "1")[:1] != "0" else "off (SPZ_DIST_PLACE=0)"}
                del big
        except Exception as e:
            line["qft_northstar"] = {"error": repr(e)}

    # ---- CPU baseline: oracle port on this box's host cores, bounded sample ----
    if rank == 0 and not args.no_cpu and world == 1:
        state = None
        import oracle as orc
        threads = orc.max_threads()
        cn = pick_cpu_n(min(n, args.cpu_qubits or 30))
        v, measured, passes, est = cpu_sweep_weighted(cn, threads)
        v1, _m1, _p1, _e1 = cpu_sweep_weighted(min(cn, 26), 1)
        line["cpu_baseline"] = {"value": v, "unit": "GB/s", "cores": threads, "kind": "port",
                                "sample": f"{passes} timed unfused passes (H/RX/RZ on the weighted target sample "
                                          f"{[(t, round(w, 2)) for t, w in cpu_targets(cn, threads)]}) of a {cn}-qubit state, {measured:.1f} s; "
                                          f"value = bytes of the full sweep / weighted seconds ({est:.1f} s)",
                                "single_thread_GBps": v1}
        if not args.no_extras:
            try:  # BASELINE config 1 (QFT-20 on the CPU, 1 thread and all host threads) next to the GPU's fused QFT-20
                line["cpu_baseline"]["qft20"] = {"one_thread": cpu_qft(20, 1), "all_threads": cpu_qft(20, threads)}
            except Exception as e:
                line["cpu_baseline"]["qft20"] = {"error": repr(e)}
            try:  # the reference's criterion suite at n = 25 on the port, all host threads (next to line["criterion"])
                cc = cpu_criterion(threads, pick_cpu_n(CRITERION_N))
                line["cpu_baseline"]["criterion"] = {"seconds": cc, "threads": threads}
                if isinstance(line.get("criterion", {}).get("gpu_seconds"), dict):
                    gs = line["criterion"]["gpu_seconds"]
                    line["criterion"]["speedup_vs_cpu_port"] = {k: round(cc[k] / gs[k], 1) for k in cc if gs.get(k)}
            except Exception as e:
                line["cpu_baseline"]["criterion"] = {"error": repr(e)}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.shutdown()


if __name__ == "__main__":
    main()
