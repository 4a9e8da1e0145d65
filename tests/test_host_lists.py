"""Host-side list builders of the engine (chalte.cu: make_group_lists, make_span_table) checked WITHOUT a GPU: the
harness in tests/native/ includes the engine's translation unit, runs the builders on real line lists / channel grids
and asserts the invariants the device kernels rely on (every active channel in exactly one group lane and one
(span, segment); a group's records and a record's line where the kernels look for them; staging-area limits)."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import helpers as H

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    out = tmp_path_factory.mktemp("native") / "host_lists_harness"
    src = os.path.join(H.ROOT, "tests", "native", "host_lists_harness.cu")
    res = subprocess.run([NVCC, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), src],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return str(out)


def _grid(mol, n_chan, v_centre):
    """Selected lines (frequency-sorted over all molecules of a joint fit 'a+b'), their molecule ids, the channel grid."""
    from cha1_mcmc_b200 import synthetic as SY
    from cha1_mcmc_b200.catalog import MolCat, find_catalog
    per = [np.sort(SY._trimmed_freqs(MolCat(m, find_catalog(SY.default_cat_folder(), m)), 7000, 30000)) for m in mol.split("+")]
    lines = np.concatenate(per); ids = np.concatenate([np.full(p.size, k) for k, p in enumerate(per)])
    order = np.argsort(lines, kind="stable")
    lines, ids = lines[order], ids[order]
    return lines, ids, SY.window_grid(lines, n_chan, SY.GOTHAM_DNU, v_centre)


@pytest.mark.parametrize("mol,n_chan,v_centre,sparse_at", [
    ("benzonitrile", 1 << 20, 0.0, 0.36),          # the headline grid (sparse): the one-pass channel stream applies
    ("1-cyanonapthalene", 1 << 14, 5.8, None),     # dense forest of lines: many segments per span
    ("hc5n_hfs", 20, 5.8, None),                   # a grid smaller than one span
    ("1-cyanonapthalene+indene_hfs", 1 << 15, 5.8, None),   # joint fit of two molecules (BASELINE config 4's pair)
])
def test_group_tile_and_span_tables_hold_their_invariants(harness, tmp_path, mol, n_chan, v_centre, sparse_at):
    lines, ids, freq = _grid(mol, n_chan, v_centre)
    lf, ff = tmp_path / "lines.bin", tmp_path / "freq.bin"
    lines.astype("<f8").tofile(lf); freq.astype("<f8").tofile(ff)
    if ids.max() > 0:
        ids.astype("<f8").tofile(str(lf) + ".mol")
    hvs = ["0.05", "0.36", "1.6", "4.0"]
    res = subprocess.run([harness, str(lf), str(ff), str(v_centre)] + hvs, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    rows = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(rows) == len(hvs)
    for r in rows:
        assert r["channels"] == freq.size and r["active"] <= freq.size and r["groups"] * 8 >= r["active"]
        if "unstaged_tiles" not in r:            # (tiles too dense to stage: no span table, the general paths run)
            assert r["segments"] >= r["nonempty_spans"] and r["fits"]
    assert "unstaged_tiles" not in rows[0] and "unstaged_tiles" not in rows[1]
    # wider windows touch more channels and more (line, channel) pairs
    assert all(a["pairs"] <= b["pairs"] and a["active"] <= b["active"] for a, b in zip(rows, rows[1:]))
    if sparse_at is not None:
        assert next(r for r in rows if abs(r["hv"] - sparse_at) < 1e-9)["sparse"]
    else:
        assert not any(r.get("sparse") for r in rows)
