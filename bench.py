#!/usr/bin/env python
"""bench.py -- GCUPS (incl. traceback) of the B200 alignment hot path, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c2|c1] [--pairs P]
    python bench.py --impl reference ...      # the reference's CPU path on the host cores

Workloads are BASELINE.json's configs.  The default is config 4
("batch of 1M short protein pairs, local alignment, sharded across 1/2/4/8 B200"):
it is the configuration the metric's 1/2/4/8-GPU quoting applies to that fits one
GPU; a "step" is one pass of the hot path over the whole batch (weak scaling: every
rank aligns its own --pairs pairs, no data-path collective).  c1/c2/c3 are the
single-pair configs (replicas only at N>1).

  value  = whole-job GCUPS with inputs resident in HBM (sa_align_batch_device /
           sa_align_device on torch's stream), fill + device traceback + string emission;
  e2e    = the same metric through the host-buffer C ABI (sa_align_batch / sa_align):
           pinned host inputs -> H2D -> kernels -> D2H of results and strings, per step.
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
sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402

BLOSUM62 = None


def load_matrix(name):
    mats = json.load(open(os.path.join(ROOT, "tests", "golden", "matrices.json")))
    return np.asarray(mats[name], np.int32)


def peaks():
    p = dict(hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            p = dict(hbm_gbs=float(json.load(open(f))["hbm_gbs"]), source="MEASURED_PEAKS.json")
        except Exception:
            pass
    # integer / DPX issue rate measured by csrc/microbench/pipe_peaks.cu on this pool's B200
    dpx = 18.25e12
    f = os.path.join(ROOT, "profiles", "r01_pipe_peaks.jsonl")
    if os.path.exists(f):
        for line in open(f):
            try:
                d = json.loads(line)
            except Exception:
                continue
            if d.get("op") == "VIADDMNMX":
                dpx = d["lane_ops_per_s_T"] * 1e12
    p["dpx_lane_ops_per_s"] = dpx
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "25"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                    power_w_max=max(power), samples=len(sm))


# ------------------------------------------------------------------------------------------ workloads
def make_workload(args, rank):
    if args.workload == "c4":
        T, toff, P, poff = synth.synthetic_batch(args.pairs, seed=2024 + rank)
        return dict(kind="batch", mode=1, alpha=23, matrix=load_matrix("protein/blosum62.txt"), gap=5,
                    text=T, toff=toff, pattern=P, poff=poff,
                    cells=int(((toff[1:] - toff[:-1] + 1) * (poff[1:] - poff[:-1] + 1)).sum()),
                    name=f"c4: {args.pairs} protein pairs/GPU, ~300 aa, mutate.py-style partners, SW, BLOSUM62, gap 5")
    seqs = np.load(os.path.join(ROOT, "tests", "golden", "sequences.npz"))
    if args.workload == "c1":
        a, b = seqs["dna/NC_018874.txt"], seqs["dna/mutated_NC_018874.txt"]
        w = dict(mode=0, alpha=4, matrix=load_matrix("dna/blast.txt"), gap=5,
                 name="c1: data/dna NC_018874 x mutated_NC_018874, NW, blast.txt, gap 5")
    elif args.workload == "c2":
        a, b = seqs["protein/P33450.fasta"], seqs["protein/mutated_P33450.fasta"]
        w = dict(mode=1, alpha=23, matrix=load_matrix("protein/blosum62.txt"), gap=5,
                 name="c2: data/protein P33450 x mutated_P33450, SW, BLOSUM62, gap 5")
    elif args.workload == "c3":
        a, b = synth.synthetic_pair(args.length, 12345, 54321)
        w = dict(mode=0, alpha=4, matrix=load_matrix("dna/blast.txt"), gap=5,
                 name=f"c3: synthetic DNA {args.length} x ~{int(args.length * 0.952)} (mutate.py-style, seeds 12345/54321), NW, blast.txt, gap 5")
    else:
        raise SystemExit("unknown workload")
    t, p = (a, b) if len(a) >= len(b) else (b, a)
    w.update(kind="single", text=np.ascontiguousarray(t), pattern=np.ascontiguousarray(p),
             cells=(len(t) + 1) * (len(p) + 1))
    return w


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation (oracle/_ref, else the oracle port) on the host cores."""
    if rank != 0:
        return
    from oracle.oracle_py import Oracle, Reference
    w = make_workload(args, 0) if args.workload != "c4" else None
    cores = os.cpu_count() or 1
    kind = "reference" if Reference.available("O3") else "port"
    if args.workload == "c4":
        sample_pairs = min(args.pairs, args.ref_pairs)
        T, toff, P, poff = synth.synthetic_batch(sample_pairs, seed=2024)
        mat = load_matrix("protein/blosum62.txt")
        cells = int(((toff[1:] - toff[:-1] + 1) * (poff[1:] - poff[:-1] + 1)).sum())
        if kind == "reference":
            ref = Reference("O3")
            step = lambda: ref.align_batch(1, 23, mat, 5, T, toff, P, poff, nthreads=cores)
        else:
            orc = Oracle()
            cores = 1
            step = lambda: [orc.align(1, 23, mat, 5, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]]) for i in range(sample_pairs)]
        sample = f"{sample_pairs} pairs of the c4 batch per step (seed 2024), alignSequenceCPU incl. traceback, one pair per thread"
        name = f"c4: {args.pairs} protein pairs/GPU, ~300 aa, mutate.py-style partners, SW, BLOSUM62, gap 5"
    else:
        t, p = w["text"], w["pattern"]
        lim = args.ref_length
        t, p = t[:lim], p[:lim]
        cells = (len(t) + 1) * (len(p) + 1)
        cores = 1
        if kind == "reference":
            ref = Reference("O3")
            step = lambda: ref.align(w["mode"], w["alpha"], w["matrix"], w["gap"], t, p)
        else:
            orc = Oracle()
            step = lambda: orc.align(w["mode"], w["alpha"], w["matrix"], w["gap"], t, p)
        sample = f"leading {len(t)} x {len(p)} sub-problem of the pair, alignSequenceCPU incl. traceback, 1 thread"
        name = w["name"]
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    gcups = cells * args.steps / dt / 1e9
    line = dict(metric="GCUPS incl. traceback", value=gcups, unit="GCUPS", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="int32", data="synthetic", impl="reference",
                config=dict(workload=name),
                cpu_baseline=dict(value=gcups, unit="GCUPS", cores=cores, kind=kind, sample=sample),
                e2e=dict(value=gcups, unit="GCUPS", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def cpu_baseline(args, w):
    from oracle.oracle_py import Oracle, Reference
    cores = os.cpu_count() or 1
    kind = "reference" if Reference.available("O3") else "port"
    if w["kind"] == "batch":
        k = min(len(w["toff"]) - 1, args.ref_pairs)
        toff, poff = w["toff"][:k + 1], w["poff"][:k + 1]
        cells = int(((toff[1:] - toff[:-1] + 1) * (poff[1:] - poff[:-1] + 1)).sum())
        if kind == "reference":
            ref = Reference("O3")
            t0 = time.perf_counter()
            ref.align_batch(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"], toff, w["pattern"], poff, nthreads=cores)
            dt = time.perf_counter() - t0
            t1 = time.perf_counter()
            k1 = max(1, k // cores)
            ref.align_batch(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"], toff[:k1 + 1], w["pattern"], poff[:k1 + 1], nthreads=1)
            dt1 = time.perf_counter() - t1
            cells1 = int(((toff[1:k1 + 1] - toff[:k1] + 1) * (poff[1:k1 + 1] - poff[:k1] + 1)).sum())
            return dict(value=cells / dt / 1e9, unit="GCUPS", cores=cores, kind=kind,
                        sample=f"first {k} pairs of the batch, unmodified alignSequenceCPU (-O3) incl. traceback, one pair per thread",
                        one_core_gcups=cells1 / dt1 / 1e9)
        orc = Oracle()
        k = min(k, 4000)
        t0 = time.perf_counter()
        for i in range(k):
            orc.align(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"][toff[i]:toff[i + 1]], w["pattern"][poff[i]:poff[i + 1]])
        dt = time.perf_counter() - t0
        cells = int(((toff[1:k + 1] - toff[:k] + 1) * (poff[1:k + 1] - poff[:k] + 1)).sum())
        return dict(value=cells / dt / 1e9, unit="GCUPS", cores=1, kind="port", sample=f"first {k} pairs, oracle/sa_oracle.c")
    t, p = w["text"][:args.ref_length], w["pattern"][:args.ref_length]
    cells = (len(t) + 1) * (len(p) + 1)
    eng = Reference("O3") if kind == "reference" else Oracle()
    t0 = time.perf_counter()
    eng.align(w["mode"], w["alpha"], w["matrix"], w["gap"], t, p)
    dt = time.perf_counter() - t0
    return dict(value=cells / dt / 1e9, unit="GCUPS", cores=1, kind=kind,
                sample=f"leading {len(t)} x {len(p)} sub-problem, alignSequenceCPU incl. traceback, 1 thread")


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    sa = load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    al = sa.Aligner(local_rank)
    w = make_workload(args, rank)
    tstream = torch.cuda.Stream(device=dev)           # the kernels are launched (and timed) on this stream
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    pk = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if w["kind"] == "batch":
        N = len(w["toff"]) - 1
        arena = int(w["toff"][-1] + w["poff"][-1])
        max_n = int((w["toff"][1:] - w["toff"][:-1]).max())
        max_m = int((w["poff"][1:] - w["poff"][:-1]).max())
        # pinned host copies (the e2e path reads these), device-resident copies (the `value` path)
        hT = torch.from_numpy(w["text"]).pin_memory(); hP = torch.from_numpy(w["pattern"]).pin_memory()
        hto = torch.from_numpy(w["toff"]).pin_memory(); hpo = torch.from_numpy(w["poff"]).pin_memory()
        dT, dP, dto, dpo = (x.to(dev, non_blocking=True) for x in (hT, hP, hto, hpo))
        dres = torch.zeros(N * 4, dtype=torch.int64, device=dev)
        daoff = torch.zeros(N, dtype=torch.int64, device=dev)
        doT = torch.empty(arena, dtype=torch.uint8, device=dev)
        doP = torch.empty(arena, dtype=torch.uint8, device=dev)
        hout = dict(results=torch.zeros(N * 4, dtype=torch.int64).pin_memory().numpy().view(sa.RESULT_DTYPE),
                    aln_off=torch.zeros(N, dtype=torch.int64).pin_memory().numpy().view(np.uint64),
                    aligned_text=torch.empty(arena, dtype=torch.uint8).pin_memory().numpy(),
                    aligned_pattern=torch.empty(arena, dtype=torch.uint8).pin_memory().numpy())

        def dev_step():
            al.align_batch_device(w["mode"], w["alpha"], w["matrix"], w["gap"], N, dT.data_ptr(), dto.data_ptr(),
                                  dP.data_ptr(), dpo.data_ptr(), dres.data_ptr(), daoff.data_ptr(), doT.data_ptr(),
                                  doP.data_ptr(), arena, max_n, max_m, stream=stream)

        def e2e_step():
            al.align_batch(w["mode"], w["alpha"], w["matrix"], w["gap"], hT.numpy(), hto.numpy(), hP.numpy(), hpo.numpy(), out=hout)

        h2d = int(hT.numel() + hP.numel() + 16 * (N + 1))
        d2h = int(32 * N + 8 * N + 2 * arena)
        l2_note = f"inputs {h2d / 1e6:.0f} MB and direction workspace >> 126 MB L2 (no flush needed)"
    else:
        n, m = len(w["text"]), len(w["pattern"])
        hT = torch.from_numpy(w["text"]).pin_memory(); hP = torch.from_numpy(w["pattern"]).pin_memory()
        dT, dP = hT.to(dev), hP.to(dev)
        doT = torch.empty(n + m, dtype=torch.uint8, device=dev)
        doP = torch.empty(n + m, dtype=torch.uint8, device=dev)
        dres = torch.zeros(4, dtype=torch.int64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def dev_step():
            flush.fill_(1)          # evict L2 between timed iterations (inputs are smaller than L2)
            al.align_device(w["mode"], w["alpha"], w["matrix"], w["gap"], dT.data_ptr(), n, dP.data_ptr(), m,
                            doT.data_ptr(), doP.data_ptr(), dres.data_ptr(), stream=stream)

        def e2e_step():
            al.align(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"], w["pattern"])

        h2d, d2h = n + m, 2 * (n + m) + 32
        l2_note = "L2 flushed between iterations (256 MiB fill)"

    # ---- device-resident timing (`value`) ----
    for _ in range(args.warmup):
        dev_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fill_us = tb_us = 0.0
    launches = 0
    ev0.record()
    per_step = []
    for _ in range(args.steps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        dev_step()
        b.record()
        per_step.append((a, b))
    ev1.record()
    barrier()
    clocks = sampler.stop()
    t = al.timing()                     # kernel events of the LAST step
    launches = t["kernel_launches"]
    # kernel durations for the roofline: one extra step with the fill/traceback overlap switched off,
    # so that the CUDA events bracket the fill kernels alone (same kernels, same inputs)
    os.environ["SA_BATCH_PIPELINE"] = "0"
    dev_step()
    barrier()
    t = al.timing()
    os.environ.pop("SA_BATCH_PIPELINE", None)
    fill_us, tb_us = t["fill_us"], t["traceback_us"]
    if w["kind"] == "single":
        step_ms = [x.elapsed_time(y) for x, y in per_step]
        # the L2 flush is not part of the path: use the kernel events of the path itself
        dev_ms = (fill_us + tb_us) / 1e3
        total_ms = dev_ms * args.steps
    else:
        total_ms = ev0.elapsed_time(ev1)
        dev_ms = total_ms / args.steps
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    cc = torch.tensor([float(w["cells"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
    total_ms_max = float(tt.item()); cells_all = float(cc.item())
    value = cells_all * args.steps / (total_ms_max * 1e-3) / 1e9

    # ---- end-to-end through the host-buffer C ABI ----
    for _ in range(max(1, args.warmup - 1)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    et = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_value = cells_all * args.steps / float(et.item()) / 1e9
    te = al.timing()
    if te.get("d2h_bytes"):      # the batch entry point reports what its last call copied (strings are packed on the device)
        h2d, d2h = int(te["h2d_bytes"]), int(te["d2h_bytes"])

    if rank == 0:
        cells = float(w["cells"])
        fill_s = fill_us * 1e-6
        gcups_fill = cells / fill_s / 1e9 if fill_s > 0 else None
        alg_bytes = 0.25 * cells + h2d        # packed 2-bit directions written + sequences read, per step
        hbm_ach = alg_bytes / fill_s / 1e9 if fill_s > 0 else None
        dpx_peak_gcups = pk["dpx_lane_ops_per_s"] / 2 / 1e9      # 2 DPX-class ops per cell in s32
        hbm_peak_gcups = pk["hbm_gbs"] / 0.25
        bind = min(dpx_peak_gcups, hbm_peak_gcups)
        per_gpu_value = value / world
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get(args.workload)
            except Exception:
                traffic = None
        line = dict(
            metric="GCUPS incl. traceback", value=value, unit="GCUPS", n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=total_ms_max / args.steps, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype="int32", data="synthetic" if args.workload in ("c3", "c4") else "reference data/ files",
            config=dict(workload=w["name"], l2=l2_note, per_gpu_cells_per_step=w["cells"],
                        parallelism=("pairs sharded over ranks, no collective" if w["kind"] == "batch" else "replicas only")),
            e2e=dict(value=e2e_value, unit="GCUPS", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     ms_per_step=float(et.item()) / args.steps * 1e3),
            gpu_launches=int(launches) * args.steps,
            clocks=clocks,
            roofline=dict(bound="hbm", achieved=hbm_ach, peak=pk["hbm_gbs"], unit="GB/s",
                          frac=(hbm_ach / pk["hbm_gbs"]) if hbm_ach else None, traffic=traffic,
                          peak_source=pk["source"], kernel="batch_line16_kernel (all class launches of the step)" if w["kind"] == "batch" else "long_fill_kernel",
                          kernel_ms_per_step=fill_us / 1e3, traceback_ms_per_step=tb_us / 1e3,
                          algorithmic_bytes_per_step=alg_bytes,
                          note="HBM roofline of the packed direction matrix (0.25 B/cell); the binding roofline is DPX-ALU, see roofline_dpx"),
            roofline_dpx=dict(bound="dpx-alu", achieved=(2 * cells / fill_s) if fill_s > 0 else None,
                              peak=pk["dpx_lane_ops_per_s"], unit="lane-ops/s (2 DPX ops per cell)",
                              frac=(2 * cells / fill_s / pk["dpx_lane_ops_per_s"]) if fill_s > 0 else None,
                              peak_source="csrc/microbench/pipe_peaks.cu on B200 (profiles/r01_pipe_peaks.jsonl)",
                              fill_only_gcups=gcups_fill, roofline_gcups=bind,
                              frac_incl_traceback=per_gpu_value / bind),
        )
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, w)
        print(json.dumps(line))
    al.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="c4: pairs per GPU")
    ap.add_argument("--length", type=int, default=100_000, help="c3: text length")
    ap.add_argument("--ref-pairs", type=int, default=40_000, help="CPU sample size (pairs) for c4")
    ap.add_argument("--ref-length", type=int, default=30_000, help="CPU sample size (residues) for single-pair workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
