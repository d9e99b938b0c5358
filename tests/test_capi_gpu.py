"""The two self-contained C-ABI entry points (include/praline_b200.h: pgpu_align_batch,
pgpu_align_profiles) called straight through ctypes -- no Engine, no Python planning -- against the
oracle.  torch only provides the device buffers."""
import ctypes

import numpy as np
import pytest

import oracle
from praline_b200 import _lib, matrices, synth
from conftest import MODES

pytestmark = pytest.mark.gpu
MODE_ID = {"global": 0, "local": 1, "semiglobal_both": 2, "semiglobal_one": 3, "semiglobal_two": 4}


@pytest.fixture(scope="module")
def lib():
    import torch
    lib = _lib.load()
    _lib.check(lib.pgpu_init(0))
    torch.cuda.set_device(0)
    return lib


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _batch_call(lib, mode, seqs, pi, pj, S, gaps, want_paths):
    import torch
    flat, offs = synth.pack(seqs)
    lens = np.asarray([len(s) for s in seqs], np.int64)
    n = len(pi)
    seqs_d, offs_d = _dev(flat.astype(np.uint8)), _dev(offs.astype(np.int64))
    pi_d, pj_d, S_d = _dev(np.asarray(pi, np.int32)), _dev(np.asarray(pj, np.int32)), _dev(S.astype(np.float32))
    scores = torch.empty(n, dtype=torch.float32, device="cuda")
    caps = lens[pi] + lens[pj] + 2
    reg = np.zeros(n + 1, np.int64)
    np.cumsum(caps, out=reg[1:])
    pbuf = torch.zeros((int(reg[-1]), 2), dtype=torch.int32, device="cuda") if want_paths else None
    plen = torch.zeros(n, dtype=torch.int32, device="cuda") if want_paths else None
    go, ge = (gaps[0], gaps[-1])
    _lib.check(lib.pgpu_align_batch(MODE_ID[mode], n, _ptr(seqs_d), _ptr(offs_d), _ptr(pi_d), _ptr(pj_d), _ptr(S_d),
                                    S.shape[0], go, ge, int(want_paths), _ptr(scores), _ptr(pbuf), _ptr(plen), None))
    torch.cuda.synchronize()
    sc = scores.cpu().numpy()
    if not want_paths:
        return sc, None
    pb, pl = pbuf.cpu().numpy(), plen.cpu().numpy()
    return sc, [pb[reg[k + 1] - pl[k]:reg[k + 1]] for k in range(n)]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("gaps", [[-11.0, -1.0], [-4.0]])
def test_align_batch_entry_point(lib, mode, gaps):
    S = matrices.blosum62()
    rng = np.random.default_rng(7)
    seqs = synth.family(3, 14, 90) + synth.family(4, 5, 333) + [rng.integers(0, 20, int(l)).astype(np.int32) for l in (1, 2, 31, 33, 700)]
    n = len(seqs)
    pi = rng.integers(0, n, 300)
    pj = rng.integers(0, n, 300)          # unsorted, repeated, i == j included
    flat, offs = synth.pack(seqs)
    got, _ = _batch_call(lib, mode, seqs, pi, pj, S, gaps, False)
    want, wpaths = oracle.align_batch(mode, flat, offs, pi, pj, S, gaps, want_paths=True)
    assert np.array_equal(got, want)
    got, paths = _batch_call(lib, mode, seqs, pi, pj, S, gaps, True)
    assert np.array_equal(got, want)
    for k in range(len(pi)):
        assert np.array_equal(paths[k], wpaths[k]), (mode, gaps, k)


def test_align_batch_errors(lib):
    import torch
    S = matrices.blosum62()
    seqs = [np.zeros(1500, np.int32), np.ones(10, np.int32)]
    flat, offs = synth.pack(seqs)
    keep = []

    def hold(t):
        keep.append(t)       # device buffers must outlive the call that gets their pointers
        return _ptr(t)

    args = lambda pi, pj, Sm, want: (0, 1, hold(_dev(flat.astype(np.uint8))), hold(_dev(offs.astype(np.int64))),
                                     hold(_dev(np.asarray([pi], np.int32))), hold(_dev(np.asarray([pj], np.int32))),
                                     hold(_dev(Sm)), 27, -11.0, -1.0, want, hold(torch.empty(1, device="cuda")),
                                     hold(torch.empty((1600, 2), dtype=torch.int32, device="cuda")),
                                     hold(torch.empty(1, dtype=torch.int32, device="cuda")), None)
    assert lib.pgpu_align_batch(*args(1, 0, S, 0)) == 4                       # sequence two beyond 1024
    assert b"1024" in lib.pgpu_last_error()
    assert lib.pgpu_align_batch(*args(0, 1, (S * 0.5).astype(np.float32), 1)) == 4   # traced needs integer scores
    assert lib.pgpu_align_batch(*args(0, 1, S, 1)) == 0                       # sequence ONE may be long
    assert lib.pgpu_align_batch(9, 1, *args(0, 1, S, 0)[2:]) == 1


@pytest.mark.parametrize("mode", MODES)
def test_align_profiles_entry_point(lib, mode):
    import torch
    S = matrices.blosum62()
    c1, c2 = synth.count_profile(11, 83, 6, 20, 27), synth.count_profile(12, 140, 5, 20, 27)
    p1, p2 = synth.profile_from_counts(c1), synth.profile_from_counts(c2)
    rng = np.random.default_rng(3)
    q1, q2 = rng.random((83, 5)).astype(np.float32), rng.random((140, 5)).astype(np.float32)   # a second track set
    S2 = rng.integers(-3, 6, (5, 5)).astype(np.float32)
    L1, L2 = 83, 140
    g1 = np.stack([rng.integers(-12, -8, L1), rng.integers(-3, 0, L1)], axis=1).astype(np.float32)
    g2 = np.stack([rng.integers(-12, -8, L2), rng.integers(-3, 0, L2)], axis=1).astype(np.float32)
    zero = [(y, x) for y in range(20, 41) for x in range(30, 77)] if mode == "local" else None
    m = oracle.build_scores([p1, q1], [p2, q2], [S, S2])
    want_score, want_path = oracle.align_raw(mode, m, g1, g2, zero_idxs=zero)
    devs = [_dev(a) for a in (p1, q1, p2, q2, S.astype(np.float32), S2, g1, g2)]
    arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    z_d = None
    if zero:
        z = np.zeros((L1 + 1, L2 + 1), np.uint8)
        z[tuple(np.asarray(zero).T)] = 1
        z_d = _dev(z)
    score = torch.zeros(1, dtype=torch.float32, device="cuda")
    path = torch.zeros((L1 + L2 + 2, 2), dtype=torch.int32, device="cuda")
    plen = torch.zeros(1, dtype=torch.int32, device="cuda")
    A = (ctypes.c_int * 2)(27, 5)
    _lib.check(lib.pgpu_align_profiles(MODE_ID[mode], 2, arr(devs[0:2]), arr(devs[2:4]), arr(devs[4:6]), A, L1, L2,
                                       _ptr(devs[6]), _ptr(devs[7]), _ptr(z_d), _ptr(score), _ptr(path), _ptr(plen), None))
    torch.cuda.synchronize()
    assert float(score.item()) == want_score
    n = int(plen.item())
    assert np.array_equal(path.cpu().numpy()[L1 + L2 + 2 - n:], want_path)


@pytest.mark.parametrize("mode", ["global", "semiglobal_both", "local"])
def test_align_profile_long_entry_point(lib, mode):
    """pgpu_align_profile_long straight through ctypes on caller-allocated buffers (padded score matrix, workspace of
    pgpu_general_workspace_bytes): score matrix beside the fill for the global / semiglobal modes (2^21 cells and
    more), in sequence for local; a non-default stream; score, end cell and path against the oracle; errors."""
    import torch
    S = matrices.nucleotide()
    L1, L2 = 1600, 1377
    p1 = synth.profile_from_counts(synth.count_profile(31, L1, 8, 4, 15))
    p2 = synth.profile_from_counts(synth.count_profile(32, L2, 8, 4, 15))
    g1, g2 = oracle.gap_arrays(L1, L2, [-11.0, -1.0])
    want_s, want_p = oracle.align_raw(mode, oracle.build_scores([p1], [p2], [S]), g1, g2)
    pitch = (L2 + 127) // 128 * 128
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d1, d2, dS, dg1, dg2 = _dev(p1), _dev(p2), _dev(S.astype(np.float32)), _dev(g1), _dev(g2)
        m = torch.empty((L1, pitch), dtype=torch.float32, device="cuda")
        ws = torch.empty(int(lib.pgpu_general_workspace_bytes(L1, L2)), dtype=torch.uint8, device="cuda")
        out = torch.zeros(8 + 2 * (L1 + L2 + 2), dtype=torch.int32, device="cuda")
        base = out.data_ptr()
        for rep in range(2):
            _lib.check(lib.pgpu_align_profile_long(MODE_ID[mode], _ptr(d1), _ptr(d2), _ptr(dS), 15, L1, L2, _ptr(m), pitch,
                                                   _ptr(dg1), _ptr(dg2), 0, _ptr(ws), ctypes.c_void_p(base),
                                                   ctypes.c_void_p(base + 4), ctypes.c_void_p(base + 32),
                                                   ctypes.c_void_p(base + 16), ctypes.c_void_p(base + 20),
                                                   ctypes.c_void_p(st.cuda_stream)))
            st.synchronize()
            h = out.cpu().numpy()
            assert float(h[:1].view(np.float32)[0]) == want_s, (mode, rep)
            got = h[8:].reshape(-1, 2)[int(h[4]):int(h[4]) + int(h[5])]
            assert np.array_equal(got, want_p), (mode, rep)
        assert np.array_equal(m[:, :L2].cpu().numpy(), oracle.build_scores([p1], [p2], [S]))     # the matrix it built
    assert lib.pgpu_align_profile_long(MODE_ID[mode], _ptr(d1), _ptr(d2), _ptr(dS), 15, 0, L2, _ptr(m), pitch, _ptr(dg1),
                                       _ptr(dg2), 0, _ptr(ws), None, None, None, None, None, None) != 0
    assert b"empty" in lib.pgpu_last_error()


def test_plain_c_program(lib, tmp_path):
    """examples/capi_demo.c: the library driven from plain C (cudart only, no Python in the data path).
    Built with gcc here, run as a subprocess; every printed score and path equals the oracle's."""
    import os
    import shutil
    import subprocess
    from conftest import ROOT
    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this box")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = str(tmp_path / "capi_demo")
    libdir = os.path.join(ROOT, "praline_b200")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                    os.path.join(ROOT, "examples", "capi_demo.c"), "-o", exe, "-L", libdir, "-lpraline_b200",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + libdir], check=True)
    S = matrices.blosum62().astype(np.float32)
    mfile = str(tmp_path / "blosum62.f32")
    S.tofile(mfile)
    n, L = 12, 40
    out = subprocess.run([exe, mfile, "27", str(n), str(L)], check=True, capture_output=True, text=True).stdout
    # the program's LCG, restated
    seed = [12345]

    def lcg():
        seed[0] = (seed[0] * 1664525 + 1013904223) & 0xffffffff
        return seed[0] >> 8
    lens = [L - 3 + lcg() % 7 for _ in range(n)]
    seqs = [np.array([lcg() % 20 for _ in range(l)], np.int32) for l in lens]
    lines = out.strip().splitlines()
    assert len(lines) == n * (n - 1) // 2
    for line in lines:
        parts = line.split()
        i, j, score = int(parts[0]), int(parts[1]), float(parts[2])
        path = np.array([[int(v) for v in p.split(",")] for p in parts[3:]], np.int32)
        want_score, want_path = oracle.align_seqs("global", seqs[i], seqs[j], S, [-11.0, -1.0])
        assert score == want_score, (i, j)
        assert np.array_equal(path, want_path), (i, j)
