"""Hyper-parameter sweep over run.py -- what the reference does with GNU parallel in
batch-job.sh:43-52 (`parallel --retry-failed --joblog logs/... -jN 'python run.py --name={1} -k={2} ...' ::: ...`):
the cartesian product of the flag values, N jobs at a time on a device, failed jobs retried, one log line per job.

cfg1-sized models (nltcs: 40 MFLOP per step) are launch-latency bound, so several runs share a GPU: `--jobs` is per
device and `--devices` lists the GPUs; every job gets its device through run.py's own `--device` flag.

    python sweep.py --name nltcs kdd -k 100 --dim 30 50 70 --batch 128 --epoch 250 --rate 0.0005 0.001 \
        --cost 0.25 0.5 1 --seed 11 --ema --devices 0 1 --jobs 4 --joblog logs/log_nltcs_kdd
"""
from __future__ import annotations

import argparse
import itertools
import os
import shlex
import subprocess
import sys
import threading
import time
from typing import Dict, List, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))

# (flag of run.py, argparse dest here) in the order of the reference's command line (batch-job.sh:44)
AXES = [("--name", "name"), ("-k", "embedding"), ("--dim", "dim"), ("--batch", "batch"), ("--epoch", "epoch"),
        ("--rate", "rate"), ("--cost", "cost"), ("--decay", "decay"), ("--seed", "seed")]


def build_grid(args) -> List[List[str]]:
    """One argument list for run.py per point of the grid (axes that were not given keep run.py's default)."""
    axes = [(flag, getattr(args, dest)) for flag, dest in AXES if getattr(args, dest)]
    jobs = []
    for combo in itertools.product(*[vals for _, vals in axes]):
        cmd: List[str] = []
        for (flag, _), v in zip(axes, combo):
            cmd += [flag, str(v)]
        if args.ema:
            cmd.append("--ema")
        if args.note:
            cmd += ["--note", args.note]
        jobs.append(cmd)
    return jobs


def run_sweep(jobs: Sequence[Sequence[str]], devices: Sequence[int], jobs_per_device: int, runner: Sequence[str],
              joblog: str = "", retries: int = 1, quiet: bool = False) -> List[Dict]:
    """Runs every job (`runner + job + ['--device', d]`), at most `jobs_per_device` at a time per device.
    Returns one record per job: seq, device, start, runtime, exitval, tries, command."""
    lock = threading.Lock()
    queue = list(enumerate(jobs, 1))
    records: Dict[int, Dict] = {}
    log = None
    if joblog:
        os.makedirs(os.path.dirname(os.path.abspath(joblog)), exist_ok=True)
        log = open(joblog, "a")
        if log.tell() == 0:
            log.write("Seq\tDevice\tStarttime\tJobRuntime\tExitval\tTries\tCommand\n")

    def worker(device: int):
        while True:
            with lock:
                if not queue:
                    return
                seq, job = queue.pop(0)
            cmd = list(runner) + list(job) + ["--device", str(device)]
            rec = {"seq": seq, "device": device, "command": " ".join(shlex.quote(c) for c in cmd), "tries": 0}
            for attempt in range(1 + max(0, retries)):          # parallel --retry-failed
                rec["tries"] = attempt + 1
                rec["start"] = time.time()
                p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                rec["runtime"] = time.time() - rec["start"]
                rec["exitval"] = p.returncode
                rec["output"] = p.stdout[-4000:]
                if p.returncode == 0:
                    break
            with lock:
                records[seq] = rec
                if log:
                    log.write(f"{seq}\t{device}\t{rec['start']:.3f}\t{rec['runtime']:.3f}\t{rec['exitval']}\t"
                              f"{rec['tries']}\t{rec['command']}\n")
                    log.flush()
                if not quiet:
                    last = rec["output"].strip().splitlines()[-1:] or [""]
                    print(f"[{len(records)}/{len(jobs)}] rc={rec['exitval']} {rec['runtime']:.1f}s dev{device}: {last[0]}",
                          flush=True)

    threads = [threading.Thread(target=worker, args=(d,), daemon=True)
               for d in devices for _ in range(max(1, jobs_per_device))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if log:
        log.close()
    return [records[k] for k in sorted(records)]


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="grid of run.py jobs, several per GPU (reference: batch-job.sh)")
    ap.add_argument("--name", "-n", nargs="+", required=True)
    ap.add_argument("--embedding", "-k", nargs="+", type=int, required=True)
    ap.add_argument("--dim", "-d", nargs="+", type=int, required=True)
    ap.add_argument("--batch", "-b", nargs="+", type=int)
    ap.add_argument("--epoch", "-e", nargs="+", type=int)
    ap.add_argument("--rate", "-r", nargs="+", type=float)
    ap.add_argument("--cost", "-c", nargs="+", type=float)
    ap.add_argument("--decay", "-g", nargs="+", type=float)
    ap.add_argument("--seed", "-s", nargs="+", type=int)
    ap.add_argument("--ema", "-m", action="store_true")
    ap.add_argument("--note", "-t", default="")
    ap.add_argument("--devices", nargs="+", type=int, default=[0])
    ap.add_argument("--jobs", "-j", type=int, default=1, help="concurrent runs per device")
    ap.add_argument("--joblog", default="")
    ap.add_argument("--retries", type=int, default=1)
    ap.add_argument("--runner", default="", help="command that replaces 'python run.py' (tests)")
    args = ap.parse_args(argv)
    runner = shlex.split(args.runner) if args.runner else [sys.executable, os.path.join(HERE, "run.py")]
    recs = run_sweep(build_grid(args), args.devices, args.jobs, runner, args.joblog, args.retries)
    failed = [r for r in recs if r["exitval"] != 0]
    print(f"{len(recs) - len(failed)} of {len(recs)} runs finished; results appended to result.txt by run.py")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
