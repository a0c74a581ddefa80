"""The C-driven ranking pass (csrc/pass.cu: pps_pass_begin / _count / _end) - several distance blocks, top-k admission in
the distance epilogue, a gallery sharded over ranks with two collectives - against the launch-by-launch Python path of
round 1 (same kernels, so every output must agree bit for bit) and, sharded, against the single-device result.

The sharded protocol is exercised ON ONE GPU: every "rank" gets its own pps_ctx and its gallery block, the all-reduce of
the thresholds and the all-gather of the packed [keys | counters | flags] buffers are done by hand between the C calls -
exactly the bytes NCCL would move (tests/test_gpu_multi.py runs the same thing over NCCL when >= 2 GPUs are visible)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pps_oracle as O

pytestmark = pytest.mark.gpu


def _engine(evaluator, d, q, g, **kw):
    return evaluator.RankEngine(d["qid"], d["gid"], d["qcam"], d["gcam"], nq=q.shape[0], ng_local=g.shape[0], dim=q.shape[1], **kw)


def _same(a, b, topk):
    np.testing.assert_array_equal(a.ap, b.ap)
    np.testing.assert_array_equal(a.is_valid, b.is_valid)
    np.testing.assert_array_equal(a.first_rank, b.first_rank)
    if topk:
        np.testing.assert_array_equal(a.topk_index, b.topk_index)
        np.testing.assert_array_equal(a.topk_dist, b.topk_dist)


@pytest.mark.parametrize("name", ["small_mid", "ragged_dim", "many_pos", "dup_ties", "some_invalid"])
@pytest.mark.parametrize("topk", [0, 13])
@pytest.mark.parametrize("dtype", ["fp32", "fp16"])
def test_pass_equals_python_path(golden, name, topk, dtype):
    import torch
    from pps_b200 import evaluator
    d = golden(name)
    tdt = torch.float16 if dtype == "fp16" else torch.float32
    q, g = torch.from_numpy(d["q"]).cuda().to(tdt), torch.from_numpy(d["g"]).cuda().to(tdt)
    for block_bytes in (8 << 30, q.shape[0] * 256 * 4):
        ref = _engine(evaluator, d, q, g, topk=topk, max_block_bytes=block_bytes, in_dtype=tdt)
        ref.use_c_pass = ref.use_c_path = False
        eng = _engine(evaluator, d, q, g, topk=topk, max_block_bytes=block_bytes, in_dtype=tdt)
        eng.use_c_path = False
        _same(eng.run(q, g), ref.run(q, g), topk)
        _same(eng.run(q, g), ref.run(q, g), topk)          # a second pass on the same context


def test_pass_topk_overflow_falls_back_to_the_sweep(golden):
    """A 4-entry candidate buffer and a gallery whose nearest rows come last: the epilogue admission overflows, the flag
    comes back with the results and the pass is repeated with the one-read sweep."""
    import torch
    from pps_b200 import evaluator
    d = golden("small_mid")
    order = np.argsort(-O.compute_dist(d["q"][:1], d["g"])[0])
    dd = dict(d, gid=d["gid"][order], gcam=d["gcam"][order])
    g = torch.from_numpy(np.ascontiguousarray(d["g"][order])).cuda()
    q = torch.from_numpy(d["q"]).cuda()
    want = _engine(evaluator, dd, q, g, topk=13).run(q, g)
    eng = _engine(evaluator, dd, q, g, topk=13, max_block_bytes=q.shape[0] * 256 * 4)
    eng.tk_cap = 4
    got = eng.run(q, g)
    assert eng.fused_topk and not eng.used_fused_topk
    _same(got, want, 13)
    eng.tk_cap = 0
    got = eng.run(q, g)
    assert eng.used_fused_topk
    _same(got, want, 13)


def _emulated_sharded_pass(torch, lib, _lib, evaluator, d, q, g_full, world, topk, block_bytes, in_code, prec):
    """world ranks on one device: own ctx + gallery block each, the two exchanges done by hand."""
    dev = q.device
    nq, dim, ng = int(q.shape[0]), int(q.shape[1]), int(g_full.shape[0])
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)
    qid, qcam, gid, gcam = up(d["qid"]), up(d["qcam"]), up(d["gid"]), up(d["gcam"])
    s = _lib.stream_ptr()
    ctxs, shards, x1s = [], [], []
    for r in range(world):
        h = C.c_void_p(0)
        _lib.check(lib.pps_ctx_create(0, C.byref(h)), "ctx")
        ctxs.append(h)
        row0, rows = evaluator.gallery_shard(ng, r, world)
        gl = g_full[row0:row0 + rows].contiguous()
        shards.append(gl)
        d_x1, n_x1 = C.c_void_p(0), C.c_longlong(0)
        _lib.check(lib.pps_pass_begin(h, _lib.ptr(q), nq, _lib.ptr(gl) if rows else None, rows, dim, in_code, _lib.ptr(qid),
                                      _lib.ptr(qcam), _lib.ptr(gid), _lib.ptr(gcam), ng, row0, world, r, prec, topk, block_bytes,
                                      0, s, C.byref(d_x1), C.byref(n_x1)), "begin")
        x1s.append(evaluator._wrap_device(torch, d_x1.value, max(n_x1.value, 1), "<i4", torch.int32, dev)[:n_x1.value])
    assert len({int(x.numel()) for x in x1s}) == 1
    g1 = torch.cat(x1s).contiguous() if x1s[0].numel() else None      # all-gather of the thresholds
    packed = []
    for h in ctxs:
        d_x2, nb = C.c_void_p(0), C.c_longlong(0)
        _lib.check(lib.pps_pass_count(h, _lib.ptr(g1), s, C.byref(d_x2), C.byref(nb)), "count")
        packed.append(evaluator._wrap_device(torch, d_x2.value, nb.value, "|u1", torch.uint8, dev))
    assert len({int(p.numel()) for p in packed}) == 1
    gathered = torch.cat(packed).contiguous()                        # all-gather
    outs = []
    for h in ctxs:
        out_map = C.c_double(0.0)
        out_cmc = np.zeros(10)
        ap, valid, first = np.zeros(nq), np.zeros(nq, np.uint8), np.zeros(nq, np.int32)
        ti = np.zeros((nq, topk), np.int32) if topk else None
        td = np.zeros((nq, topk), np.float32) if topk else None
        rc = lib.pps_pass_end(h, _lib.ptr(gathered), 10, s, C.cast(C.byref(out_map), C.c_void_p), _lib.ptr(out_cmc), _lib.ptr(ap),
                              _lib.ptr(valid), _lib.ptr(first), _lib.ptr(ti), _lib.ptr(td))
        assert rc in (0, _lib.PPS_ERR_NO_VALID_QUERY), rc
        outs.append(evaluator.RankResult(ap, valid, first, None, None, ti, td))
    for h in ctxs:
        lib.pps_ctx_destroy(h)
    return outs


@pytest.mark.parametrize("name,world", [("small_mid", 2), ("dup_ties", 3), ("many_pos", 2), ("some_invalid", 4)])
@pytest.mark.parametrize("topk", [0, 12])
def test_sharded_pass_emulated_on_one_gpu_equals_single_device(golden, name, world, topk):
    import torch
    from pps_b200 import _lib, evaluator
    lib = _lib.load()
    d = golden(name)
    q, g = torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda()
    prec = _lib.PRECISIONS[evaluator.DEFAULT_PRECISION]
    one = _engine(evaluator, d, q, g, topk=topk).run(q, g)
    for block_bytes in (8 << 30, q.shape[0] * 256 * 4):
        for res in _emulated_sharded_pass(torch, lib, _lib, evaluator, d, q, g, world, topk, block_bytes, _lib.DTYPE_F32, prec):
            _same(res, one, topk)


def test_sharded_pass_with_an_empty_shard(golden):
    """More ranks than 256-row blocks make sense for: one rank holds no gallery row at all, and still takes part in both
    exchanges with buffers of the same size."""
    import torch
    from pps_b200 import _lib, evaluator
    lib = _lib.load()
    d = golden("small_mid")
    q = torch.from_numpy(d["q"]).cuda()
    keep = 3                                                           # 3 gallery rows over 4 ranks
    dd = dict(d, gid=d["gid"][:keep], gcam=d["gcam"][:keep])
    g = torch.from_numpy(d["g"][:keep]).cuda()
    one = _engine(evaluator, dd, q, g, topk=2).run(q, g)
    for res in _emulated_sharded_pass(torch, lib, _lib, evaluator, dd, q, g, 4, 2, 8 << 30, _lib.DTYPE_F32,
                                      _lib.PRECISIONS[evaluator.DEFAULT_PRECISION]):
        _same(res, one, 2)


def test_pass_argument_errors():
    import torch
    from pps_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p(0)
    _lib.check(lib.pps_ctx_create(0, C.byref(h)), "ctx")
    s = _lib.stream_ptr()
    t = torch.zeros(64, device="cuda")
    ids = torch.zeros(8, dtype=torch.int64, device="cuda")
    bad = lambda **kw: lib.pps_pass_begin(h, _lib.ptr(t), kw.get("nq", 1), _lib.ptr(t), kw.get("ngl", 1), 64, kw.get("dtype", 0),
                                          _lib.ptr(ids), _lib.ptr(ids), _lib.ptr(ids), _lib.ptr(ids), kw.get("ng", 1),
                                          kw.get("off", 0), kw.get("world", 1), kw.get("rank", 0), kw.get("prec", _lib.PREC_F16X3),
                                          kw.get("topk", 0), kw.get("bb", 1 << 20), 0, s, None, None)
    assert bad(nq=0) == _lib.PPS_ERR_INVALID_ARG
    assert bad(off=1) == _lib.PPS_ERR_INVALID_ARG            # offset + local rows > global rows
    assert bad(rank=1) == _lib.PPS_ERR_INVALID_ARG
    assert bad(topk=1000) == _lib.PPS_ERR_INVALID_ARG
    assert bad(prec=_lib.PREC_FP32) == _lib.PPS_ERR_INVALID_ARG
    assert lib.pps_pass_count(h, None, s, None, None) == _lib.PPS_ERR_INVALID_ARG      # no pass in flight
    assert lib.pps_pass_end(h, None, 10, s, None, None, None, None, None, None, None) == _lib.PPS_ERR_INVALID_ARG
    assert bad() == 0
    lib.pps_ctx_destroy(h)


@pytest.mark.parametrize("dtype", ["fp32", "fp16"])
def test_pass_host_input_equals_resident(golden, dtype):
    """RankEngine.run_host: pinned host rows -> staging buffers filled by pps_pass_begin (slabs of a one-block shard,
    whole blocks otherwise) == the pass on resident tensors, bit for bit; repeated calls reuse the staging."""
    import torch
    from pps_b200 import evaluator
    d = golden("many_pos")
    tdt = torch.float16 if dtype == "fp16" else torch.float32
    qh, gh = torch.from_numpy(d["q"]).to(tdt).pin_memory(), torch.from_numpy(d["g"]).to(tdt).pin_memory()
    q, g = qh.cuda(), gh.cuda()
    for block_bytes in (8 << 30, q.shape[0] * 256 * 4):
        for topk in (0, 9):
            eng = _engine(evaluator, d, q, g, topk=topk, max_block_bytes=block_bytes, in_dtype=tdt)
            eng.use_c_path = False
            want = eng.run(q, g)
            _same(eng.run_host(qh, gh), want, topk)
            _same(eng.run_host(qh, gh), want, topk)
    # a large single block: several slabs
    from pps_b200 import synthetic
    dd = synthetic.make_reid_set(nq=300, ng=20000, dim=256, n_ids=60, n_cams=4, n_distractors=3000, sigma=3.0, seed=5)
    qh, gh = torch.from_numpy(dd["q"]).to(tdt).pin_memory(), torch.from_numpy(dd["g"]).to(tdt).pin_memory()
    q, g = qh.cuda(), gh.cuda()
    eng = _engine(evaluator, dd, q, g, topk=5, in_dtype=tdt)
    eng.use_c_path = False
    _same(eng.run_host(qh, gh), eng.run(q, g), 5)
    with pytest.raises(RuntimeError, match="run_host"):
        eng.run_host(q, gh)


def test_numa_helper_is_best_effort():
    from pps_b200 import numa
    info = numa.bind_to_gpu_node(0)
    assert info["device"] == 0 and ("reason" in info or info["bound"])
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}


def test_speculative_pass_resizes_when_the_ids_change():
    """The second pass of a shape lays its buffers out for the sizes of the first (no mid-pass read-back).  When the ids
    change under it - more pairs, longer lists, more candidate rows - the bound check fails, pps_pass_end says
    PPS_ERR_PASS_RESIZE, and the engine repeats the pass as a sizing pass: results as if nothing had happened."""
    import torch
    from pps_b200 import evaluator, synthetic
    rs = np.random.RandomState(3)
    nq, ng, dim = 300, 40000, 128
    q = torch.from_numpy(rs.randn(nq, dim).astype(np.float32)).cuda()
    g = torch.from_numpy(rs.randn(ng, dim).astype(np.float32)).cuda()
    qcam, gcam = rs.randint(0, 3, nq), rs.randint(0, 3, ng)
    qid = rs.randint(1, 60, nq)
    few = np.where(rs.rand(ng) < 0.02, rs.randint(1, 60, ng), 0)          # ~2 % of the rows carry a query id
    many = np.where(rs.rand(ng) < 0.30, rs.randint(1, 60, ng), 0)         # 15x more pairs, candidates, list lengths
    for topk, bb in ((0, 8 << 30), (7, nq * 4096 * 4)):
        outs = {}
        for name, gid in (("few", few), ("many", many), ("few_again", few)):
            eng = evaluator.RankEngine(qid, gid, qcam, gcam, nq=nq, ng_local=ng, dim=dim, topk=topk, max_block_bytes=bb)
            eng.use_c_path = False
            a = eng.run(q, g)
            b = eng.run(q, g)                                             # speculative on the sizes of the pass before
            ref = evaluator.RankEngine(qid, gid, qcam, gcam, nq=nq, ng_local=ng, dim=dim, topk=topk, max_block_bytes=bb)
            ref.use_c_pass = ref.use_c_path = False
            want = ref.run(q, g)
            _same(a, want, topk)
            _same(b, want, topk)
            outs[name] = a
        assert not np.array_equal(outs["few"].ap, outs["many"].ap)
        np.testing.assert_array_equal(outs["few"].ap, outs["few_again"].ap)


@pytest.mark.parametrize("name", ["small_mid", "dup_ties", "some_invalid"])
def test_single_call_path_equals_pass(golden, name):
    """pps_evaluate_device_ctx (round 1's one-call path for one device / one block / no top-k, `use_c_path`) and the pass
    produce the same bits."""
    import torch
    from pps_b200 import evaluator
    d = golden(name)
    q, g = torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda()
    one = _engine(evaluator, d, q, g)
    one.use_c_path = True
    two = _engine(evaluator, d, q, g)
    assert not two.use_c_path and two.use_c_pass
    _same(one.run(q, g), two.run(q, g), 0)


@pytest.mark.parametrize("name", ["small_mid", "dup_ties", "some_invalid", "many_pos"])
@pytest.mark.parametrize("topk", [0, 13])
def test_counting_epilogue_equals_block_and_count_kernel(golden, name, topk):
    """fp16 rows (single-plane operands), several blocks: the blocks after the first take the counting epilogue with top-k
    admission (pps_dist_rank_topk_tc: no distance block is written); PPS_PASS_NO_FUSED_COUNT writes and counts them.  Same
    bits, on a first (sizing) pass and on the speculative passes after it."""
    import torch
    from pps_b200 import evaluator
    d = golden(name)
    q, g = torch.from_numpy(d["q"]).cuda().half(), torch.from_numpy(d["g"]).cuda().half()
    bb = q.shape[0] * 256 * 4
    ref = _engine(evaluator, d, q, g, topk=topk, max_block_bytes=bb, in_dtype=torch.float16)
    ref.fused_count = False
    want = ref.run(q, g)
    assert not ref.used_fused_count
    eng = _engine(evaluator, d, q, g, topk=topk, max_block_bytes=bb, in_dtype=torch.float16)
    for _ in range(3):
        got = eng.run(q, g)
        _same(got, want, topk)
    per_query = np.array([(d["gid"] == i).sum() for i in d["qid"]])
    # the epilogue tables hold <= 64 thresholds per query (speculative passes plan for 8 more than the last sizing pass saw)
    assert eng.used_fused_count == (eng.pass_blocks > 1 and 0 < per_query.max() <= 56), (eng.pass_blocks, per_query.max())
