"""Host logic of the sweep driver (the reference's batch-job.sh:43-52 grid): grid construction, jobs per device,
retry of failed jobs, job log.  The runner is a stub -- no GPU involved."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pgm-vae_b200"))

import sweep  # noqa: E402


def test_grid_is_the_cartesian_product_in_reference_flag_order():
    ap_args = ["--name", "kdd", "nltcs", "-k", "4096", "--dim", "10", "--batch", "32", "--epoch", "200", "--rate", "0.0002",
               "--cost", "0.35", "0.4", "0.45", "0.5", "--seed", "5", "--note", "50_40_30_20"]
    import argparse
    ns = argparse.Namespace(name=["kdd", "nltcs"], embedding=[4096], dim=[10], batch=[32], epoch=[200], rate=[0.0002],
                            cost=[0.35, 0.4, 0.45, 0.5], decay=None, seed=[5], ema=False, note="50_40_30_20")
    jobs = sweep.build_grid(ns)
    assert len(jobs) == 8 and len(ap_args) > 0
    assert jobs[0] == ["--name", "kdd", "-k", "4096", "--dim", "10", "--batch", "32", "--epoch", "200", "--rate", "0.0002",
                       "--cost", "0.35", "--seed", "5", "--note", "50_40_30_20"]
    assert jobs[-1][1] == "nltcs" and jobs[-1][13] == "0.5"


def test_jobs_run_on_their_devices_failed_ones_are_retried(tmp_path):
    stub = tmp_path / "stub.py"
    marker = tmp_path / "seen"
    # fails the first time it sees a given seed, succeeds on the retry; records the device it was given
    stub.write_text(
        "import sys, os\n"
        "a = sys.argv[1:]\n"
        "seed, dev = a[a.index('--seed') + 1], a[a.index('--device') + 1]\n"
        f"p = os.path.join({str(tmp_path)!r}, 'try_' + seed)\n"
        "first = not os.path.exists(p)\n"
        "open(p, 'a').write(dev + '\\n')\n"
        f"open({str(marker)!r}, 'a').write(seed + ' ' + dev + '\\n')\n"
        "print('seed', seed, 'device', dev)\n"
        "sys.exit(3 if first and seed == '2' else 0)\n")
    jobs = [["--name", "nltcs", "--seed", str(s)] for s in range(6)]
    log = tmp_path / "logs" / "joblog"
    recs = sweep.run_sweep(jobs, devices=[0, 1], jobs_per_device=2, runner=[sys.executable, str(stub)], joblog=str(log),
                           retries=1, quiet=True)
    assert [r["seq"] for r in recs] == [1, 2, 3, 4, 5, 6]
    assert all(r["exitval"] == 0 for r in recs)
    assert [r["tries"] for r in recs] == [1, 1, 2, 1, 1, 1]
    assert {r["device"] for r in recs} <= {0, 1}
    lines = log.read_text().strip().splitlines()
    assert lines[0].startswith("Seq\tDevice") and len(lines) == 7
    # the device in the command line is the device of the record
    for r in recs:
        assert r["command"].endswith(f"--device {r['device']}")


def test_cli_reports_failures(tmp_path):
    stub = tmp_path / "fail.py"
    stub.write_text("import sys; sys.exit(1)\n")
    rc = sweep.main(["--name", "nltcs", "-k", "8", "--dim", "2", "--runner", f"{sys.executable} {stub}", "--retries", "0"])
    assert rc == 1
