"""CPU tests of the host side: structure matrices and spin operators pinned to the reference's, the
TensorNetwork container, the scheduler, and the C-ABI surface (loads, exports every declared symbol,
fails loudly without a GPU).  Mirrors the reference's unit tests tests/test_main.py:68-103."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_structure_matrices_match_reference(tnsu):
    g = load_golden("structure_matrices")
    for key in g.files:
        if key.startswith("inf_"):
            mine = tnsu.smc.infinite_structure_matrix_dict(key[4:])
        elif key.startswith("pbc_"):
            mine = tnsu.smc.square_peps_pbc(int(key[4:]))
        elif key.startswith("obc_"):
            h, w = key[4:].split("x")
            mine = tnsu.smc.rectangular_peps_obc(int(h), int(w))
        else:
            continue
        assert mine.shape == g[key].shape and np.array_equal(mine, g[key]), key
        assert tnsu.smc.is_valid(mine)


def test_spin_operators(tnsu):
    """reference tests/test_main.py:68-91 plus the golden matrices up to spin 2."""
    g = load_golden("structure_matrices")
    for s in (0.5, 1.0, 1.5, 2.0):
        assert np.abs(np.stack(tnsu.spin_operators(s)) - g[f"spin_{s}"]).max() < 1e-15
    sx, sy, sz = tnsu.spin_operators(0.5)
    assert np.sum(np.abs(sx - np.array([[0, 0.5], [0.5, 0.0]]))) == 0
    assert np.sum(np.abs(sy - np.array([[0, -0.5j], [0.5j, 0.0]]))) == 0
    assert np.sum(np.abs(sz - np.array([[0.5, 0.0], [0.0, -0.5]]))) == 0
    sx1, sy1, sz1 = tnsu.spin_operators(1.0)
    assert np.sum(np.abs(sx1 - np.array([[0, 1.0, 0.0], [1.0, 0.0, 1.0], [0, 1.0, 0.0]]) / np.sqrt(2))) < 1e-12
    assert np.sum(np.abs(sy1 - np.array([[0, -1j, 0.0], [1j, 0.0, -1j], [0, 1j, 0.0]]) / np.sqrt(2))) < 1e-12
    assert np.sum(np.abs(sz1 - np.diag([1.0, 0.0, -1.0]))) < 1e-12
    with pytest.raises(AssertionError):
        tnsu.spin_operators(0.3)


def test_structure_matrix_validation(tnsu, capsys):
    """reference tests/test_main.py:94-103"""
    smc = tnsu.smc
    assert not smc.is_valid(np.array([[[1.0]]]))
    assert not smc.is_valid(np.array([[1, 2, 3, 3, 0, 0], [0, 4, 0, 2, 1, 3], [3, 0, 1, 0, 4, 2]]))
    assert not smc.is_valid(np.array([[1, 2, 3, 6, 0, 0], [0, 6, 0, 2, 1, 3], [3, 0, 1, 0, 6, 2]]))
    assert not smc.is_valid(np.array([[1, 2, 3, 0, 4, 0], [0, 4, 0, 2, 1, 3], [3, 0, 1, 0, 4, 2]]))
    assert not smc.is_valid(np.inf)
    capsys.readouterr()


def test_tensor_network_container(tnsu, tmp_path, capsys):
    smat = tnsu.smc.infinite_structure_matrix_dict("star")
    np.random.seed(3)
    tn = tnsu.TensorNetwork(structure_matrix=smat, spin_dim=2, virtual_dim=3)
    # same stream consumption as the reference: one normal draw per tensor in row order
    np.random.seed(3)
    for t in tn.tensors:
        assert np.array_equal(t, np.random.normal(loc=np.ones(t.shape), scale=1.0))
    assert all(np.allclose(w, 1 / 3) and len(w) == 3 for w in tn.weights)
    assert tn.is_valid()
    with pytest.raises(AssertionError):
        tnsu.TensorNetwork(structure_matrix=np.array([[1, 2], [1, 0]]))
    with pytest.raises(AssertionError):
        tnsu.TensorNetwork(structure_matrix=smat, spin_dim=0)
    with pytest.raises(AssertionError):
        tnsu.TensorNetwork(structure_matrix=smat, random_init_real_loc=None)
    with pytest.raises(AssertionError):  # weight length mismatch
        tnsu.TensorNetwork(structure_matrix=smat, tensors=tn.tensors, weights=[np.ones(2)] * 9)
    cx = tnsu.TensorNetwork(structure_matrix=smat, random_init_imag_loc=0.0, random_init_imag_scale=1.0)
    assert np.iscomplexobj(cx.tensors[0])
    # pickle round trip in the reference's state_dict format
    tn.dir_path, tn.network_name = str(tmp_path), "net"
    tn.save_network()
    other = tnsu.TensorNetwork(structure_matrix=smat, virtual_dim=3, dir_path=str(tmp_path), network_name="net")
    other.load_network()
    assert all(np.array_equal(a, b) for a, b in zip(other.tensors, tn.tensors))
    assert set(tn.state_dict) == {"tensors", "weights", "structure_matrix", "path", "spin_dim", "virtual_size",
                                  "network_name", "su_logger"}
    # sqrt-weight helpers are inverse to each other
    st = tn.get_tensor_network_state()
    st.absorb_all_inverse_weights()
    assert all(np.allclose(a, b) for a, b in zip(st.tensors, tn.tensors))
    capsys.readouterr()


def test_library_exports_every_declared_symbol():
    from tnsu_b200 import _lib

    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("tnsu_b200.h", "tnsu_b200_debug.h"))
    declared = set(re.findall(r"\b(tnsu_[a-z0-9_]+)\s*\(", header))
    declared.discard("tnsu_engine")
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().tnsu_abi_version() == _lib.ABI_VERSION


def _cuda_present():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


@pytest.mark.skipif(_cuda_present(), reason="checks the no-GPU failure path")
def test_compute_fails_loudly_without_gpu(tnsu):
    smat = tnsu.smc.infinite_structure_matrix_dict("peps")
    tn = tnsu.TensorNetwork(structure_matrix=smat, virtual_dim=2)
    sx, sy, sz = tnsu.spin_operators(0.5)
    su = tnsu.SimpleUpdate(tn, [1.0] * 8, 0.0, [sx, sy, sz], [sx, sy, sz], [], [0.1], d_max=2, print_process=False)
    for call in (su.energy_per_site, su.simple_update, su.run, lambda: su.tensor_rdm(0)):
        with pytest.raises(tnsu.EngineError, match="no CUDA device"):
            call()


def test_schedule_levels_preserve_reference_order(tnsu):
    """Level-major execution must be a reordering of 0..m-1 that never swaps two edges sharing a tensor."""
    from tnsu_b200.engine import schedule_levels

    expected_levels = {"peps": 6, "star": 7, "triangle": 15}
    mats = {nm: tnsu.smc.infinite_structure_matrix_dict(nm) for nm in
            ["chain", "peps", "star", "cube", "triangle", "kagome_3", "kagome_12", "pyrochlore"]}
    mats["obc16"] = tnsu.smc.rectangular_peps_obc(16, 16)
    mats["pbc4"] = tnsu.smc.square_peps_pbc(4)
    expected_levels["obc16"] = 60
    for nm, smat in mats.items():
        lv = schedule_levels(smat)
        if nm in expected_levels:
            assert lv.max() + 1 == expected_levels[nm], nm
        ends = [set(np.nonzero(smat[:, e])[0]) for e in range(smat.shape[1])]
        for a in range(smat.shape[1]):
            for b in range(a + 1, smat.shape[1]):
                if ends[a] & ends[b]:
                    assert lv[a] < lv[b], (nm, a, b)  # conflicting edges keep their sequential order
        for e in range(smat.shape[1]):  # levels are tight: some conflicting predecessor sits one level below
            if lv[e] > 0:
                assert any(ends[a] & ends[e] and lv[a] == lv[e] - 1 for a in range(e))
    with pytest.raises(tnsu.EngineError):
        schedule_levels(np.array([[1, 2], [1, 0]]))


def test_partition_round_robin():
    from tnsu_b200.sharding import gather_points, partition

    for n, w in ((256, 8), (10, 4), (3, 8), (7, 1)):
        parts = [partition(n, w, r) for r in range(w)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        partition(4, 2, 2)
    vals = np.arange(7.0)
    assert np.array_equal(gather_points(vals, 7, 1, 0)[:, 0], vals)


def test_stacked_expm_is_bit_identical_to_single_calls():
    """SimpleUpdateBatch._gates exponentiates all distinct (H, dt) pairs in one stacked scipy.linalg.expm call; the
    reference calls expm once per edge (simple_update.py:470-473).  Same Pade routine per matrix: identical bits."""
    from scipy import linalg

    pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
    hams = np.array([(-np.kron(pz, pz) + h * (np.kron(px, np.eye(2)) + np.kron(np.eye(2), px)) / 4).astype(complex)
                     for h in -np.linspace(0, 4, 33)])
    rng = np.random.default_rng(0)
    herm = rng.normal(size=(7, 9, 9)) + 1j * rng.normal(size=(7, 9, 9))
    herm = herm + np.conj(np.transpose(herm, (0, 2, 1)))
    for stack in (hams, herm):
        for dt in (0.1, 0.001, 0.05j):
            assert np.array_equal(linalg.expm(-dt * stack), np.array([linalg.expm(-dt * h) for h in stack]))


def test_schedule_gates_are_deduplicated_once_and_bit_identical():
    """_set_schedule_gates searches the distinct (Hamiltonian, dt schedule) rows ONCE for all time steps and
    exponentiates only those; every (point, edge, step) gate must equal the reference's one-matrix call
    expm(-dt H) (simple_update.py:470-473) bit for bit — shared and per-point schedules, repeated Hamiltonians, -0.0."""
    from scipy import linalg
    from tnsu_b200.simple_update import SimpleUpdateBatch

    pz, px = np.array([[1.0, 0], [0, -1.0]]), np.array([[0, 1.0], [1.0, 0]])
    fields = [-0.5, -2.0, -0.5, 0.0, -0.0, -3.0]  # points 0 / 2 and 3 / 4 share their Hamiltonians
    n_p, m = len(fields), 3
    hams = np.empty((n_p, m, 4, 4), dtype=complex)
    for p, h in enumerate(fields):
        for e in range(m):
            hams[p, e] = -np.kron(pz, pz) + h * (np.kron(px, np.eye(2)) / (2 + e % 2) + np.kron(np.eye(2), px) / 4)
    for dts in (np.tile([0.1, 0.01, 0.001], (n_p, 1)).astype(complex),
                np.array([[0.1, 0.01, 0.001], [0.2, 0.01, 0.001], [0.1, 0.01, 0.001], [0.1, 0.05j, 0.001], [0.1, 0.05j, 0.001],
                          [0.3, 0.02, 0.002]], dtype=complex)):
        first, which = SimpleUpdateBatch._distinct_gate_inputs(hams, dts)
        assert len(first) < n_p * m  # something was shared
        for s in range(dts.shape[1]):
            gates = SimpleUpdateBatch._expm_distinct(hams, dts[:, s], first, which)
            assert gates.shape == (n_p, m, 4, 4)
            for p in range(n_p):
                dt = dts[p, s].real if np.all(dts[:, s].imag == 0) else dts[p, s]
                for e in range(m):
                    assert np.array_equal(gates[p, e], linalg.expm(-dt * hams[p, e]))


def test_run_logs_sequence():
    """SimpleUpdateBatch.run() hands back the per-point logger dicts as a lazy sequence: list semantics (len, index,
    negative index, slice, iteration), the reference's logger keys per point, arrays for whole-batch reads."""
    from tnsu_b200.simple_update import _RunLogs

    n, cap = 5, 4
    res = dict(n_checks=np.array([2, 0, 4, 1, 3], dtype=np.int32),
               log_error=np.arange(n * cap, dtype=float).reshape(n, cap) / 1024.0,
               log_energy=-np.arange(n * cap, dtype=float).reshape(n, cap),
               log_slot=np.tile(np.array([0, 0, 1, 1], dtype=np.int32), (n, 1)),
               log_step=np.tile(np.array([3, 5, 7, 9], dtype=np.int32), (n, 1)),
               converged=np.array([1, 0, 1, 0, 1], dtype=np.int32), sweeps=np.array([5, 2, 9, 3, 7], dtype=np.int32),
               final_slot=np.array([1, 0, 1, 0, 1], dtype=np.int32))
    logs = _RunLogs(res, [[0.1, 0.01]] * n)
    assert len(logs) == n and [lg["sweeps"] for lg in logs] == [5, 2, 9, 3, 7]
    assert logs[0] is logs[0] and logs[-1]["sweeps"] == 7 and [lg["sweeps"] for lg in logs[1:3]] == [2, 9]
    assert logs[2]["error"] == [8 / 1024, 9 / 1024, 10 / 1024, 11 / 1024] and logs[2]["dt"] == [0.1, 0.1, 0.01, 0.01]
    assert logs[2]["iteration"] == [3, 5, 7, 9] and logs[2]["energy"] == [-8.0, -9.0, -10.0, -11.0]
    assert logs[1]["error"] == [] and logs[1]["converged"] is False and logs[1]["final_dt"] == 0.1
    assert logs[4]["converged"] is True and logs[4]["final_dt"] == 0.01
    assert np.array_equal(logs.sweeps, res["sweeps"]) and logs.converged.tolist() == [True, False, True, False, True]
    with pytest.raises(IndexError):
        logs[n]
