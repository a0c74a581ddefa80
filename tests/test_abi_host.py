"""CPU-side checks: the C-ABI library loads and exports every symbol include/pps_b200.h declares,
its host-only entry points work without a GPU, the host layer mirrors the reference interface, and
the product refuses to run without CUDA (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import pps_oracle as O
from pps_b200 import _lib, evaluator, pooling, synthetic


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pps_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert lib.pps_abi_version() == _lib.ABI_VERSION
    assert lib.pps_strerror(_lib.PPS_ERR_NO_VALID_QUERY) == b"No valid query"
    assert lib.pps_kpad(2048) == 2048 and lib.pps_kpad(100) == 128 and lib.pps_kpad(8064) == 8064
    assert lib.pps_split_bytes(10, 100, 2) == 10 * 128 * 2 * 2


def test_pair_lists_host_code_matches_numpy():
    d = synthetic.make_reid_set(nq=70, ng=600, dim=8, n_ids=20, n_cams=3, n_distractors=50, seed=5)
    p = evaluator.PairLists(d["qid"], d["qcam"], d["gid"], d["gcam"])
    assert p.nq == 70 and p.ng == 600
    for i in range(p.nq):
        same = np.nonzero(d["gid"] == d["qid"][i])[0]
        e0, e1 = p.off[i], p.off[i + 1]
        np.testing.assert_array_equal(p.g[e0:e1], same)                       # ascending gallery index
        np.testing.assert_array_equal(p.q[e0:e1], np.full(len(same), i))
        np.testing.assert_array_equal(p.pos[e0:e1], (d["gcam"][same] != d["qcam"][i]).astype(np.uint8))
        j0, j1 = p.junk_off[i], p.junk_off[i + 1]
        np.testing.assert_array_equal(p.junk_g[j0:j1], same[d["gcam"][same] == d["qcam"][i]])
    assert p.max_pairs == int(np.diff(p.off).max())
    # junk == exactly the complement of the reference's valid mask
    i = 3
    valid = O.valid_mask(d["qid"][i], d["qcam"][i], d["gid"], d["gcam"])
    np.testing.assert_array_equal(np.nonzero(~valid)[0], p.junk_g[p.junk_off[i]:p.junk_off[i + 1]])


def test_pair_lists_empty_and_no_match():
    p = evaluator.PairLists(np.array([5, 6]), np.array([0, 1]), np.array([1, 2, 3]), np.array([0, 0, 0]))
    assert p.n_pairs == 0 and p.max_pairs == 0 and list(p.off) == [0, 0, 0]
    p = evaluator.PairLists(np.zeros(0, np.int64), np.zeros(0, np.int64), np.array([1]), np.array([0]))
    assert p.n_pairs == 0 and list(p.off) == [0]


def test_split_helper_mirrors_reference_tables():
    for n in (5, 6, 7, 9, 10):
        for scale in (1.0 / 16, 1.0 / 8):
            assert pooling.uniform_partition_split(n, 384, scale) == O.uniform_partition_split(n, 384, scale)
    assert pooling.uniform_partition_split(6, 256, 1.0 / 16) == [2] * 6


def test_blob_names_and_masks():
    names = pooling.blob_names(6)
    assert len(names) == 63 and names[0] == "pps0_pool2" and names[2] == "pps01_pool2" and names[-1] == "pps012345_pool2"
    masks = [pooling.comb_to_mask(c) for c in pooling.pyramid_combs]
    assert len(masks) == 21 and masks[0] == 1 and masks[-1] == 63
    assert pooling.mask_to_comb(0b101001, 6) == [0, 3, 5]


def test_rank_result_aggregation_matches_oracle(golden):
    """RankResult turns per-query kernel outputs into the reference's averages; feed it the oracle's
    count-based outputs and compare with the oracle's sort-based cmc / mean_ap."""
    d = golden("small_mid")
    dist = d["dist"]
    ap, valid, first, neg_before = O.rank_counts(dist, d["qid"], d["gid"], d["qcam"], d["gcam"])
    p = evaluator.PairLists(d["qid"], d["qcam"], d["gid"], d["gcam"])
    nb = np.zeros(max(p.n_pairs, 1), dtype=np.int32)
    for i in range(p.nq):
        e = np.arange(p.off[i], p.off[i + 1])
        nb[e[p.pos[e] == 1]] = neg_before[i]
    res = evaluator.RankResult(ap, valid, first, nb, p)
    ids = dict(query_ids=d["qid"], gallery_ids=d["gid"], query_cams=d["qcam"], gallery_cams=d["gcam"])
    assert abs(res.mean_ap() - float(d["mAP"])) < 1e-12
    np.testing.assert_allclose(res.cmc(10, True), d["cmc_fmb"], atol=1e-12)
    np.testing.assert_allclose(res.cmc(20, False), d["cmc_all"], atol=1e-12)
    rows, v = res.cmc(10, True, average=False)
    np.testing.assert_array_equal(rows, d["cmc_rows"])


def test_no_valid_query_error():
    res = evaluator.RankResult(np.zeros(3), np.zeros(3, np.uint8), np.full(3, -1, np.int32))
    with pytest.raises(RuntimeError, match="No valid query"):
        res.cmc(10, True)


def test_reid_results_dict_shape():
    r = evaluator.reid_results((0.5, np.arange(10) / 10.0, None, None), "market")
    assert list(r.keys()) == ["market"]
    assert r["market"]["ReID"]["mAP"] == 0.5 and r["market"]["ReID"]["CMC5"] == 0.4 and r["market"]["ReID"]["mq_mAP"] == -1


def test_parse_im_name():
    assert evaluator.parse_im_name("00000012_0003_00000007.jpg", "id") == 12
    assert evaluator.parse_im_name("00000012_0003_00000007.jpg", "cam") == 3


def test_product_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pooling.pps_pool(torch.zeros(1, 32, 24, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        evaluator.compute_dist(np.zeros((2, 8), np.float32), np.zeros((3, 8), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_json_dataset_reader_orders_by_image_id_and_reads_marks(tmp_path):
    import json
    from pps_b200 import dataset_io
    images = [dict(id=3, file_name="00000007_0001_00000002.jpg"), dict(id=1, file_name="00000005_0000_00000000.jpg"),
              dict(id=2, file_name="00000005_0002_00000001.jpg")]
    anns = [dict(id=10, image_id=1, mark=0), dict(id=11, image_id=2, mark=1), dict(id=12, image_id=3, mark=2)]
    path = tmp_path / "split.json"
    path.write_text(json.dumps(dict(images=images, annotations=anns)))
    ds = dataset_io.JsonReidDataset(str(path))
    roidb = ds.get_roidb(gt=True)
    assert [e["id"] for e in roidb] == [1, 2, 3] and [e["mark"] for e in roidb] == [0, 1, 2]
    assert evaluator.get_info(roidb[2])[:2] == (7, 1)
    path.write_text(json.dumps(dict(images=images, annotations=anns[:2])))
    with pytest.raises(RuntimeError, match="no annotation"):
        dataset_io.JsonReidDataset(str(path))


def test_plan_blocks_covers_the_gallery_exactly():
    """Host logic of the multi-block pass: blocks tile [0, ng) without gaps, the first one is short only when the
    top-k bound has to be established, and a gallery that fits one block is one block."""
    assert evaluator.plan_blocks(0, 1000) == []
    assert evaluator.plan_blocks(700, 1000, topk=100) == [(0, 700)]
    assert evaluator.plan_blocks(2500, 1000) == [(0, 1000), (1000, 1000), (2000, 500)]
    for ng, blk, k in ((10_000_000, 637_440, 100), (1_250_000, 637_440, 100), (2_000_000, 637_440, 128), (900, 256, 7)):
        blocks = evaluator.plan_blocks(ng, blk, topk=k)
        assert blocks[0][0] == 0 and sum(r for _, r in blocks) == ng
        assert all(a[0] + a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert all(0 < r <= blk for _, r in blocks)
        if blk >= 65536:
            first = blocks[0][1]
            assert first % 256 == 0 and 32768 <= first < blk
            assert k * blk / first <= 2048 / 2.0                      # expected admissions per query and block
        else:
            assert blocks[0][1] == blk                                # tiny blocks (tests): nothing to shorten


def _build_c_example(tmp_path):
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    lib = _lib.lib_path() if os.path.exists(_lib.lib_path()) else None
    _lib.load()
    libdir = os.path.dirname(_lib.lib_path())
    exe = os.path.join(str(tmp_path), "evaluate_host")
    cmd = [gcc, "-O2", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "evaluate_host.c"),
           "-L", libdir, "-lpps_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 0, proc.stdout
    return exe


def test_header_is_plain_c_and_the_example_links(tmp_path):
    """include/pps_b200.h compiles as C99 with -Wall -Werror, examples/evaluate_host.c links against the shared library
    through the C ABI alone; without a CUDA device the call fails loudly (no CPU path)."""
    import subprocess
    exe = _build_c_example(tmp_path)
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if not has_cuda:
        proc = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert proc.returncode != 0 and "pps_evaluate_host" in proc.stdout


@pytest.mark.gpu
def test_c_example_runs_on_the_device(tmp_path):
    import subprocess
    exe = _build_c_example(tmp_path)
    proc = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert proc.returncode == 0, proc.stdout
    first = proc.stdout.splitlines()[0]
    assert first.startswith("ABI %d" % _lib.ABI_VERSION) and "mAP" in first
    m_ap = float(first.split("mAP")[1].split()[0])
    assert 0.05 < m_ap <= 1.0


def test_host_partition_helpers_properties():
    """Property tests (hypothesis) of the host-side partition logic: gallery shards and gallery blocks tile their range
    exactly, in order, for any sizes; the strip tables always cover the map height."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 5_000_000), st.integers(1, 16))
    def shards(ng, world):
        parts = [evaluator.gallery_shard(ng, r, world) for r in range(world)]
        assert parts[0][0] == 0 and sum(rows for _, rows in parts) == ng
        assert all(a[0] + a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [rows for _, rows in parts]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)      # np.array_split rule

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 20_000_000), st.integers(256, 2_000_000), st.integers(0, 128), st.sampled_from([4, 1024, 2048, 8192]))
    def blocks(ng, blk, k, cap):
        out = evaluator.plan_blocks(ng, blk, topk=k, cand_cap=cap)
        assert sum(rows for _, rows in out) == ng
        assert all(a[0] + a[1] == b[0] for a, b in zip(out, out[1:]))
        assert all(0 < rows <= blk for _, rows in out)
        assert (not out) or out[0][0] == 0

    @settings(max_examples=100, deadline=None)
    @given(st.sampled_from([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12]), st.sampled_from([1 / 16., 1 / 8., 1 / 32.]))
    def strips(n, scale):
        split = pooling.uniform_partition_split(n, 384, scale)
        assert len(split) == n and all(s == int(s) for s in split)
        h = int(384 * scale)
        tabled = n in (5, 7, 9, 10)                      # bpm_heads.py:25-40: rows of the 24-row tables times 16 * scale
        if (tabled and scale in (1 / 16., 1 / 8.)) or (not tabled and h % n == 0):
            assert sum(split) == h                       # otherwise Caffe2's Split refuses, as the reference's graph would
        assert split == list(O.uniform_partition_split(n, 384, scale))

    shards()
    blocks()
    strips()


def test_tile_schedules_visit_every_tile_exactly_once():
    """The persistent tile walks of the 2-CTA distance kernel (replayed on the host by pps_test_tile_walk of the TEST-ONLY
    library csrc/test_hooks.cu - the product ABI exports no instrumentation): over all CTA pairs every (group, m, n) tile
    appears exactly once; in the counting-epilogue order a pair's runs keep one m tile, their n tiles are consecutive
    and run_start / run_end bracket them."""
    import ctypes
    from pps_b200 import build as _build
    hooks = ctypes.CDLL(_build.build_test_hooks())
    hooks.pps_test_tile_walk.restype = ctypes.c_longlong
    hooks.pps_test_tile_walk.argtypes = [ctypes.c_int] * 4 + [ctypes.c_longlong] * 2 + [ctypes.c_void_p, ctypes.c_longlong]
    assert not hasattr(_lib.load(), "pps_debug_tile_walk") and not hasattr(_lib.load(), "pps_test_tile_walk")
    rs = np.random.RandomState(0)
    shapes = [(14, 78, 1, 74), (14, 2490, 1, 74), (1, 1, 1, 1), (3, 5, 1, 74), (100, 7, 1, 74), (75, 3, 1, 74), (1, 500, 1, 8),
              (91, 63, 1, 74)] + [(int(rs.randint(1, 200)), int(rs.randint(1, 300)), 1, int(rs.randint(1, 80))) for _ in range(40)]
    shapes += [(3, 1, 63, 74), (91, 1, 63, 74), (2, 1, 7, 5)]                    # grouped (embedding head): plain order only
    for m_tiles, n_tiles, groups, npairs in shapes:
        for rank_order in ((0,) if groups > 1 else (0, 1)):
            seen = {}
            total = m_tiles * n_tiles * groups
            buf = np.zeros((total + 1, 3), dtype=np.int32)
            for pair in range(npairs):
                cnt = int(hooks.pps_test_tile_walk(rank_order, m_tiles, n_tiles, groups, npairs, pair, _lib.ptr(buf), total + 1))
                assert 0 <= cnt <= total
                prev = None
                for i in range(cnt):
                    grp, m, packed = int(buf[i, 0]), int(buf[i, 1]), int(buf[i, 2]) & 0xffffffff
                    n, start, end = packed & 0x3fffffff, bool(packed & 0x40000000), bool(packed & 0x80000000)
                    assert 0 <= grp < groups and 0 <= m < m_tiles and 0 <= n < n_tiles
                    key = (grp, m, n)
                    assert key not in seen, (m_tiles, n_tiles, npairs, rank_order, key)
                    seen[key] = pair
                    if rank_order:
                        if prev is None or prev[2]:
                            assert start
                        else:
                            assert not start and m == prev[0] and n == prev[1] + 1
                        prev = (m, n, end)
                if rank_order and cnt:
                    assert prev[2]                                              # the last tile closes its run
            assert len(seen) == total, (m_tiles, n_tiles, groups, npairs, rank_order, len(seen), total)
            if rank_order and npairs >= m_tiles:
                # balanced: no pair carries more than the mean + one tile per run (round 1: 498 against 471 at 14 x 2 490 x 74)
                per_pair = np.bincount(np.array(list(seen.values())), minlength=npairs)
                assert per_pair.max() <= total / npairs + m_tiles / max(npairs % m_tiles, 1) + 2, (m_tiles, n_tiles, npairs, per_pair.max())
