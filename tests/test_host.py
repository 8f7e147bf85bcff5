"""Host-side logic (CPU only): DotMap shim, minimal HDF5 reader, Channels shim, parameter catalogue,
C-ABI surface of the built library, and the multi-rank sharding / gather (gloo, world_size 2)."""
import copy
import ctypes as C
import os
import pickle
import re

import numpy as np
import pytest
import torch

from conftest import REPO, have_reference
from score_based_channels_b200 import _lib, dist as sdist
from score_based_channels_b200 import dotmap_shim, hdf5_min, params, program

REF_MAT = "/root/reference/sample_data/CDL-C_Nt64_Nr16_ULA0.50_seed4321.mat"
LOCAL_MAT = os.path.join(REPO, "fixtures_local", "CDL-C_Nt64_Nr16_ULA0.50_seed4321.mat")
MAT = REF_MAT if os.path.exists(REF_MAT) else LOCAL_MAT
REF_CKPT = "/root/reference/pretrained_models/score-deepest-cdl-c.pt"
LOCAL_CKPT = os.path.join(REPO, "fixtures_local", "score-deepest-cdl-c.pt")
CKPT = REF_CKPT if os.path.exists(REF_CKPT) else LOCAL_CKPT


def test_dotmap_shim_semantics():
    D = dotmap_shim.DotMap
    c = D()
    c.model.ngf = 8
    assert c.model.ngf == 8 and c["model"]["ngf"] == 8
    assert not c.data.logit_transform            # missing key -> empty, falsy child (ncsnv2.py:201-202,270)
    assert "logit_transform" in c.data           # ... and it is stored, like the real class
    c2 = copy.deepcopy(c)
    c2.model.ngf = 16
    assert c.model.ngf == 8
    dotmap_shim.install()
    c3 = pickle.loads(pickle.dumps(c))
    assert c3.model.ngf == 8 and c3.toDict()["model"]["ngf"] == 8


@pytest.mark.skipif(not os.path.exists(CKPT), reason="reference checkpoint not available")
def test_reference_checkpoint_loads_into_drop_in_model():
    from score_based_channels_b200 import entry_common as ec
    from score_based_channels_b200.ncsnv2 import NCSNv2Deepest
    contents = ec.load_checkpoint(CKPT)
    cfg = contents["config"]
    assert cfg.model.ngf == 8 and cfg.model.num_classes == 2311 and cfg.sampling.steps_each == 3
    m = NCSNv2Deepest(cfg)
    res = m.load_state_dict(contents["model_state"])
    assert not res.missing_keys and not res.unexpected_keys
    assert set(m.state_dict().keys()) == set(contents["model_state"].keys())
    assert torch.equal(m.sigmas, contents["model_state"]["sigmas"])
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 2, 64, 16), torch.zeros(1, dtype=torch.long))      # no CPU fallback


@pytest.mark.skipif(not have_reference(), reason="reference not mounted")
def test_param_catalogue_matches_reference_module():
    import sys
    sys.path.insert(0, "/root/reference")
    dotmap_shim.install()
    from ncsnv2.models.ncsnv2 import NCSNv2Deepest as Ref
    from score_based_channels_b200.models import make_config
    for ngf in (8, 16):
        cfg = make_config(ngf=ngf)
        cfg.device = "cpu"
        ref = {k: tuple(v.shape) for k, v in Ref(cfg).state_dict().items()}
        assert ref == dict(params.param_shapes(ngf))


@pytest.mark.skipif(not os.path.exists(MAT), reason="sample data not available")
def test_hdf5_reader_and_channels_shim(monkeypatch):
    m = hdf5_min.loadmat_v73(MAT)
    h = m["output_h"]
    assert h.shape == (100, 10, 16, 64) and np.iscomplexobj(h)          # loaders.py:29-33 view of the file
    assert abs(np.std(h[:, 0].astype(np.complex64)) - 0.363263) < 1e-5   # SURVEY.md section 8, deviation 10
    from score_based_channels_b200 import loaders
    from score_based_channels_b200.models import make_config
    monkeypatch.setenv("SBC_DATA_DIR", os.path.dirname(MAT))
    monkeypatch.setattr(loaders, "_SEARCH", ("./data", os.path.dirname(MAT)))
    cfg = make_config()
    cfg.data.channel, cfg.data.spacing_list, cfg.data.num_pilots, cfg.data.noise_std = "CDL-C", [0.5], 38, 0.01
    np.random.seed(0)
    ds = loaders.Channels(4321, cfg, norm="global")
    assert len(ds) == 100 and ds.channels.shape == (100, 16, 64) and abs(ds.std - 0.363263) < 1e-5
    assert ds.pilots.shape == (100, 64, 38) and np.allclose(np.abs(ds.pilots), 1.0)
    it = ds[3]
    assert it["H_herm"].shape == (2, 64, 16) and it["P"].shape == (64, 38) and it["P"].dtype == np.complex64
    hh = it["H_herm"][0] + 1j * it["H_herm"][1]
    assert np.allclose(hh, np.conj(ds.channels[3].T) / ds.std, atol=1e-6)
    assert np.allclose(it["Y"], ds.channels[3] @ ds.pilots[3], atol=0.2)


def test_c_abi_library_exports_every_declared_symbol():
    """The library loads and exports exactly the entry points include/sbc.h declares (no compute calls)."""
    from score_based_channels_b200 import _lib
    hdr = open(os.path.join(REPO, "include", "sbc.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(sbc_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    L.sbc_version.restype = C.c_int
    assert L.sbc_version() == 210
    L.sbc_model_create.restype = C.c_int
    h = C.c_void_p()
    assert L.sbc_model_create(None, 0, C.byref(h)) == -1        # SBC_E_ARG, never a crash
    L.sbc_last_error.restype = C.c_char_p
    assert b"null" in L.sbc_last_error()


def test_program_tiling_covers_other_geometries():
    sd = params.random_state(8, seed=3)
    for (H, W) in ((64, 16), (32, 8), (128, 32), (24, 40)):
        for prec in program.PRECISIONS:
            p = program.build_program(sd, 8, H, W, precision=prec)
            assert p.conv_flops == 51740672 * (H * W) // 1024
            assert all(op.w_len <= p.max_w_len and op.w_off % 4 == 0 for op in p.ops)
            assert all(op.w_len == 0 or (op.wbuf >= 0 and op.wbuf % 4 == 0 and op.wbuf + op.w_len <= p.arena_floats)
                       for op in p.ops)
    p = program.build_program(sd, 8, 64, 16)
    assert p.park_floats > 0 and p.smem_bytes() <= program.smem_budget_bytes(2), \
        "the shipped geometry must be planned for two resident CTAs per SM (half of the B200 shared memory each)"
    p1 = program.build_program(sd, 8, 64, 16, park=False)
    assert p1.park_floats == 0 and p1.arena_floats * 4 + 256 <= 227 * 1024, "the single-CTA plan must fit one SM"
    with pytest.raises(ValueError, match="multiples of 8"):
        program.build_program(sd, 8, 60, 16)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = sdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    total, steps = 37, 5                                  # ragged: 19 + 18

    def fake_kernel(lo, hi):                              # stands in for the per-rank sampler launch
        ids = torch.arange(lo, hi, dtype=torch.float32)
        return torch.arange(steps, dtype=torch.float32)[:, None] * 1000 + ids[None, :]

    full = sdist.run_sharded(fake_kernel, total)
    expect = torch.arange(steps, dtype=torch.float32)[:, None] * 1000 + torch.arange(total, dtype=torch.float32)[None, :]
    ok = torch.equal(full, expect)
    lo, hi = sdist.shard_range(total, rank, world)
    torch.save({"ok": ok, "lo": lo, "hi": hi}, os.path.join(tmp, "r%d.pt" % rank))
    torch.distributed.destroy_process_group()


def test_sharding_and_nmse_gather_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % i)) for i in range(2)]
    assert r[0]["ok"] and r[1]["ok"]
    assert (r[0]["lo"], r[0]["hi"], r[1]["lo"], r[1]["hi"]) == (0, 19, 19, 37)
    # contiguous, disjoint, complete for awkward sizes too
    for total in (0, 1, 7, 8, 100, 1700, 20400):
        for world in (1, 2, 3, 4, 8):
            spans = [sdist.shard_range(total, k, world) for k in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _gather_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sdist.init_from_env("gloo")
    total = 7                                             # ragged: 4 + 3
    lo, hi = sdist.shard_range(total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None, None] * torch.ones(1, 3, 2)
    full = sdist.gather_rows(local, total)
    expect = torch.arange(total, dtype=torch.float32)[:, None, None] * torch.ones(1, 3, 2)
    torch.save({"ok": torch.equal(full, expect)}, os.path.join(tmp, "g%d.pt" % rank))
    torch.distributed.destroy_process_group()


def test_gather_rows_world_size_2_gloo(tmp_path):
    """The final-estimate gather of the approximate-MMSE entry point (test_mmse.py), ragged shards."""
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 200
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert torch.load(os.path.join(str(tmp_path), "g%d.pt" % r))["ok"]


def test_oracle_dc_boost_and_early_stop_semantics():
    """Oracle side of the approximate-MMSE knobs (reference test_mmse.py:25,246 and :173,260-263): boost 1 / no stop is
    the plain loop; a stopped sample equals a run truncated after stop+1 steps; the boost scales the DC term."""
    from oracle import oracle as orc
    from score_based_channels_b200 import synth
    sd = params.random_state(8, seed=1)
    net = orc.OracleNet(sd, 8, 64, 16)
    B = 2
    P, H = synth.qpsk_pilots(B, 64, 38), synth.cdl_like_channels(B, 64, 16)
    Y = synth.received_pilots(P, H, 1.0)
    X0 = synth.cn01((B, 64, 16), np.random.default_rng(1))
    kw = dict(noise_var=1.0, alpha_step=3e-11, beta=0.01, sigma_end=2.6e-4, level_begin=0, steps_each=3, seed=3)
    Xa, la = net.ald(P, Y, X0, H, level_end=2, **kw)
    Xb, lb = net.ald(P, Y, X0, H, level_end=2, dc_boost=1.0, stop_step=99, **kw)
    assert np.array_equal(Xa, Xb) and np.array_equal(la, lb)
    Xc, lc = net.ald(P, Y, X0, H, level_end=2, stop_step=np.array([2, 5], np.int32), **kw)     # sample 0: steps 0..2
    Xd, ld = net.ald(P, Y, X0, H, level_end=1, **kw)                                            # one level = 3 steps
    assert np.array_equal(Xc[0], Xd[0]) and np.array_equal(lc[:3, 0], ld[:, 0]) and np.all(lc[3:, 0] == 0)
    assert np.array_equal(Xc[1], Xa[1])
    Xe, _ = net.ald(P, Y, X0, H, level_end=2, dc_boost=4.0, **kw)
    assert np.abs(Xe - Xa).max() > 1e-4 * np.abs(Xa).max()


def _library_plan(sd, ngf, H, W, nthreads, prec, park):
    """Engine-1 plan built by the C++ planner inside the library (host-only entry point: no GPU needed)."""
    import ctypes as C
    L = _lib.lib()
    keep, ents = [], (_lib.StateEntry * len(sd))()
    for i, (k, v) in enumerate(sd.items()):
        a = np.ascontiguousarray(v, dtype=np.float32)
        shp = np.asarray(a.shape if a.ndim else (1,), dtype=np.int64)
        keep += [a, shp]
        ents[i] = _lib.StateEntry(k.encode(), a.ctypes.data, shp.ctypes.data, len(shp))
    h, v = C.c_void_p(), _lib.Plan1View()
    _lib.check(L.sbc_plan1_build(ents, len(sd), ngf, H, W, 2, nthreads, _lib.PREC_CODE[prec], park, C.byref(h), C.byref(v)),
               "sbc_plan1_build")
    tab = np.ctypeslib.as_array(C.cast(v.op_table, C.POINTER(C.c_int32)), (v.n_ops, 32)).copy()
    geo = np.ctypeslib.as_array(C.cast(v.geo_table, C.POINTER(C.c_int32)), (v.n_geo, 8)).copy()
    blob = np.ctypeslib.as_array(C.cast(v.blob, C.POINTER(C.c_float)), (v.blob_floats,)).copy()
    meta = {f: getattr(v, f) for f in ("arena_floats", "in_off", "out_off", "post_off", "max_w_len", "park_floats",
                                       "nthreads", "conv_flops")}
    L.sbc_plan1_free(h)
    return tab, geo, blob, meta


@pytest.mark.parametrize("ngf,H,W,prec,park", [(8, 64, 16, "tf32x3", -1), (8, 64, 16, "tf32", -1), (8, 64, 16, "tf32x3", 0),
                                               (8, 32, 8, "tf32x3", -1), (8, 24, 40, "tf32x3", -1), (16, 64, 16, "tf32x3", -1),
                                               (8, 128, 32, "tf32x3", -1)])
def test_library_planner_matches_python_planner_word_for_word(ngf, H, W, prec, park):
    """The self-contained C ABI of engine 1: csrc/sbc1_plan.h must reproduce program.py exactly -- op table, geometry
    table, parameter blob (bit patterns) and every plan scalar -- so that everything the CPU suite proves about the
    Python plan (simulator vs the reference modules, thread emulation) holds for models created without Python."""
    sd = params.random_state(ngf, seed=3)
    p = program.build_program(sd, ngf, H, W, nthreads=256, precision=prec, park=None if park < 0 else bool(park))
    tab, geo, blob, meta = _library_plan(sd, ngf, H, W, 256, prec, park)
    assert np.array_equal(p.op_table(), tab)
    assert np.array_equal(p.geo_table()[:len(p.geos)], geo)
    assert blob.size == p.blob.size and np.array_equal(blob.view(np.uint32), p.blob.view(np.uint32))
    assert meta == dict(arena_floats=p.arena_floats, in_off=p.in_off, out_off=p.out_off, post_off=p.post_off,
                        max_w_len=p.max_w_len, park_floats=p.park_floats, nthreads=p.nthreads, conv_flops=p.conv_flops)


def test_library_planner_rejects_bad_input():
    sd = params.random_state(8, seed=3)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        _library_plan(sd, 8, 60, 16, 256, "tf32x3", -1)
    bad = {k: v for k, v in sd.items() if k != "res1.0.conv1.weight"}
    with pytest.raises(RuntimeError, match="res1.0.conv1.weight"):
        _library_plan(bad, 8, 64, 16, 256, "tf32x3", -1)


REF_H5 = "/root/reference/sample_data/CDL-A_Nt64_Nr16_ULA0.50.h5"


@pytest.mark.skipif(not os.path.exists(REF_H5), reason="reference sample data not mounted")
def test_minimal_hdf5_reader_on_the_shipped_cdl_a_file():
    """The second on-disk format the reference ships (sample_data/CDL-A_Nt64_Nr16_ULA0.50.h5: chunked, compound complex):
    read with the bundled minimal HDF5 reader (no h5py / hdf5storage in the image).  QPSK pilots must come out EXACT
    (+-1/sqrt(2) in both parts) -- any mis-decoded chunk, stride or compound member breaks that -- and the channels finite
    with the expected power; two element probes pin byte order and axis order."""
    d = hdf5_min.loadmat_v73(REF_H5)
    assert set(d.keys()) == {"H", "P"}
    H, P = np.ascontiguousarray(d["H"]), np.ascontiguousarray(d["P"])
    assert H.shape == (64, 16, 200) and P.shape == (64, 64, 200) and H.dtype == np.complex128 and P.dtype == np.complex128
    r = np.float64(1.0) / np.sqrt(np.float64(2.0))
    assert np.array_equal(np.abs(P.real), np.full(P.shape, r)) and np.array_equal(np.abs(P.imag), np.full(P.shape, r))
    assert np.isfinite(H.real).all() and np.isfinite(H.imag).all()
    assert abs(float(np.mean(np.abs(H) ** 2)) - 0.12251568411307497) < 1e-12
    assert H[0, 0, 0] == complex(-0.26163689087463343, -0.11179212128717307)
    assert H[63, 15, 199] == complex(0.24750735074003632, -0.0525550311589591)
    assert P[0, 0, 0] == complex(r, r) and P[63, 63, 199] == complex(-r, -r)


def test_automatic_engine_choice_follows_the_two_cta_plan():
    """NCSNv2Deepest(precision=None) picks engine 1 exactly where its planner produces the two-CTAs-per-SM plan (host-only
    query through the library's C++ planner) and the tcgen05 engine elsewhere (measured crossover)."""
    from score_based_channels_b200 import engine
    for ngf, Nt, Nr, want in ((8, 64, 16, True), (8, 32, 8, True), (8, 24, 40, True), (8, 128, 32, False), (16, 64, 16, False)):
        assert engine.engine1_runs_two_ctas_per_sm(params.random_state(ngf, seed=1), ngf, Nt, Nr) is want, (ngf, Nt, Nr)
