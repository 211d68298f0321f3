"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol
declared in include/itrails_b200.h, the reference-mirroring Python functions behave
like the reference's, and the multi-GPU plumbing works under gloo (world size 2)."""
import ctypes
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import hmm_oracle as ho
from conftest import ROOT, golden


def test_library_exports_every_declared_symbol():
    from itrails_b200 import _lib
    header = open(os.path.join(ROOT, "include", "itrails_b200.h")).read()
    declared = set(re.findall(r"\b(itr_[a-z0-9_]+)\s*\(", header))
    declared -= {"itr_ctx", "itr_status", "itr_phase"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} missing from libitrails_b200.so"
    assert _lib.load().itr_version() >= 1000
    assert _lib.load().itr_num_states(3, 3) == 27 and _lib.load().itr_num_states(5, 5) == 70


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, never compute on CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import itrails_b200 as itb
    with pytest.raises(itb.ItrailsCudaError):
        itb.Engine(0)
    m = golden("model_1_1_example.npz")
    with pytest.raises(itb.ItrailsCudaError):
        itb.loglik_wrapper(m["a"], m["b"], m["pi"], [np.zeros(10, dtype=np.int64)])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "itrails_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower().replace("# oracle", ""), f"{f} mentions the oracle"


def test_symbols_match_reference_fixture():
    import itrails_b200 as itb
    g = golden("symbols.npz")
    assert itb.get_obs_state_dct() == list(g["names"])
    off, vals = g["order_offsets"], g["order_values"]
    for s in range(625):
        assert np.array_equal(itb.get_idx_state(s), vals[off[s]:off[s + 1]])


def test_host_emission_and_viterbi_tables_equal_oracle_bitwise():
    from itrails_b200.optimizer import emission_table, viterbi_tables
    m = golden("model_2_2_example.npz")
    a, b, pi = m["a"], m["b"], m["pi"]
    assert np.array_equal(emission_table(b), ho.emission_table(b))
    V_lst = [np.array([3, 9, 300]), np.array([624])]
    for x, y in zip(viterbi_tables(a, b, pi, V_lst), ho.viterbi_tables(a, b, pi, V_lst)):
        assert np.array_equal(x, y)


MAF = """##maf version=1
a score=1
s hg38.chr1     100 8 + 1000 ACGTNacg
s panTro5.chr1  200 8 + 2000 ACGT-ACG
s gorGor5.chr1  300 8 - 3000 AC-TAACG
s ponAbe2.chr2  400 8 + 4000 CCGTAACG
s extra.chr9    1   8 + 10   GGGGGGGG

a score=2
s hg38.chr1     120 3 + 1000 AAA
s panTro5.chr1  220 3 + 2000 CCC
s gorGor5.chr1  320 3 + 3000 TTT

a score=3
s panTro5.chr1  230 2 + 2000 G--T
s hg38.chr1     130 4 - 1000 GGTT
s ponAbe2.chr2  430 4 + 4000 NNNN
s gorGor5.chr1  330 4 + 3000 ACGT
"""


def _ref_style_parse(text, sp_lst):
    """Straight restatement of read_data.py:94-117 on MAF text (Biopython-free)."""
    names = ho.obs_state_names()
    out = []
    for blk in text.split("\n\n"):
        rows = [l.split() for l in blk.splitlines() if l.startswith("s ")]
        if not rows:
            continue
        dct = {}
        for r in rows:
            sp = r[1].split(".")[0]
            if sp in sp_lst:
                dct[sp] = r[6].replace("-", "N")
        if len(dct) == 4:
            n = len(rows[-1][6])
            out.append(np.array([names.index("".join(dct[s][i] for s in sp_lst).upper())
                                 for i in range(n)], dtype=np.int64))
    return out


def test_maf_parser_and_coordinates(tmp_path):
    import itrails_b200 as itb
    p = tmp_path / "x.maf"
    p.write_text(MAF)
    sp = ["hg38", "panTro5", "gorGor5", "ponAbe2"]
    got = itb.maf_parser(str(p), sp)
    want = _ref_style_parse(MAF, sp)
    assert len(got) == 2 == len(want)
    for g_, w in zip(got, want):
        assert g_.dtype == np.int64 and np.array_equal(g_, w)
    # species order matters (column string is built in sp_lst order)
    got2 = itb.maf_parser(str(p), sp[::-1])
    assert not np.array_equal(got2[0], got[0])
    # coordinates: + strand counts up from start, gaps -> -9; - strand counts down from srcSize-start
    co = itb.parse_coordinates(str(p), sp, "panTro5")
    assert co[0] == [200, 201, 202, 203, -9, 204, 205, 206]
    assert co[1] == [230, -9, -9, 231]
    co = itb.parse_coordinates(str(p), sp, "hg38")
    assert co[1] == [870, 869, 868, 867]
    co = itb.parse_coordinates(str(p), sp, "nosuch")
    assert co[0] == [-9] * 8
    bad = tmp_path / "bad.maf"
    bad.write_text(MAF.replace("CCGTAACG", "CCGTAARG"))
    with pytest.raises(ValueError):
        itb.maf_parser(str(bad), sp)


def test_native_maf_reader_equals_restatement(tmp_path):
    """C++ reader (csrc/maf_reader.cpp) vs the NumPy restatement of read_data.py:94-220
    (oracle/maf_oracle.py) on a generated MAF with lower case, gaps, N, missing and
    duplicated species, extra species, i/e/q lines, comments and both strands."""
    import itrails_b200 as itb
    import maf_oracle as mo
    rng = np.random.default_rng(11)
    sp = ["hg38", "panTro5", "gorGor5", "ponAbe2"]
    extra = ["macFas5", "calJac3"]
    lines = ["##maf version=1 scoring=none", "# generated"]
    for blk in range(300):
        L = int(rng.integers(1, 200))
        lines.append(f"a score={blk}.0")
        rows = list(sp) + [e for e in extra if rng.random() < 0.5]
        if rng.random() < 0.15:
            rows.remove(rows[int(rng.integers(0, 4))])          # a missing species
        if rng.random() < 0.1:
            rows.append(sp[int(rng.integers(0, 4))])            # a duplicated species (last wins)
        rng.shuffle(rows)
        for name in rows:
            seq = "".join(rng.choice(list("ACGTacgtN-"), p=[.2, .2, .2, .2, .03, .03, .03, .03, .03, .05], size=L))
            strand = "+" if rng.random() < 0.7 else "-"
            start = int(rng.integers(0, 10_000))
            lines.append(f"s {name}.chr{int(rng.integers(1, 5))} {start} {L - seq.count('-')} {strand} 100000 {seq}")
            if rng.random() < 0.2:
                lines.append(f"i {name}.chr1 N 0 C 0")
        if rng.random() < 0.2:
            lines.append("e mm9.chr1 100 20 + 5000 I")
        lines.append("")
    p = tmp_path / "gen.maf"
    p.write_text("\n".join(lines) + "\n")
    want = mo.maf_parser(str(p), sp)
    for threads in (1, 0):
        sym, off, coord, coff = itb.read_data.read_maf(str(p), sp, ref="panTro5", n_threads=threads)
        assert len(off) - 1 == len(want) > 100
        for i, w in enumerate(want):
            assert np.array_equal(sym[off[i]:off[i + 1]], w)
        wc = mo.parse_coordinates(str(p), sp, "panTro5")
        assert len(coff) - 1 == len(wc)
        for i, w in enumerate(wc):
            assert coord[coff[i]:coff[i + 1]].tolist() == w
    assert itb.parse_coordinates(str(p), sp, "macFas5") == mo.parse_coordinates(str(p), sp, "macFas5")
    with pytest.raises(FileNotFoundError):
        itb.maf_parser(str(tmp_path / "missing.maf"), sp)
    empty = tmp_path / "empty.maf"
    empty.write_text("")
    assert itb.maf_parser(str(empty), sp) == []


def test_cutpoints_match_scipy():
    from scipy.stats import expon, truncexpon
    import itrails_b200 as itb
    for n, t, c in ((3, 0.8, 1.0), (5, 2.5, 0.4), (3, 30.0, 1.0), (4, 20.0, 4.0), (3, 800.0, 0.7)):
        q = np.arange(n + 1) / n
        np.testing.assert_allclose(itb.cutpoints_AB(n, t, c), truncexpon.ppf(q, b=t * c, scale=1 / c), rtol=1e-14, atol=1e-16)
        np.testing.assert_allclose(itb.cutpoints_ABC(n, c), expon.ppf(q, scale=1 / c), rtol=1e-14)
    assert itb.get_times([0.0, 1.0, 4.0], [0, 1, 2]) == [1.0, 3.0]


def test_derive_times_cases():
    """optimizer.py:419-541: every allowed time parametrisation."""
    from itrails_b200.optimizer import derive_times
    base = dict(t_2=0.4, t_upper=7.0, N_ABC=0.5, N_AB=0.5, r=1.0, n_int_AB=3, n_int_ABC=3)
    tail = -np.log(1 - 2 / 3) * 0.5 + 7.0 + 1.0
    d = derive_times(dict(base, t_1=2.0), ["t_1"])
    assert (d["t_A"], d["t_B"], d["t_C"]) == (2.0, 2.0, 2.4) and "t_1" not in d
    assert d["t_out"] == pytest.approx(2.0 + 0.4 + tail, rel=1e-15)
    d = derive_times(dict(base, t_1=2.0, t_A=1.5), ["t_1", "t_A"])
    assert (d["t_A"], d["t_B"], d["t_C"]) == (1.5, 2.0, 2.4)
    d = derive_times(dict(base, t_1=2.0, t_C=3.0), ["t_1", "t_C"])
    assert (d["t_A"], d["t_B"], d["t_C"]) == (2.0, 2.0, 3.0)
    d = derive_times(dict(base, t_A=1.0, t_B=2.0), ["t_A", "t_B"])
    assert d["t_C"] == pytest.approx(1.9)
    assert d["t_out"] == pytest.approx(((1.5 + 0.4) + 1.9) / 2 + tail)
    d = derive_times(dict(base, t_A=1.0, t_C=2.0), ["t_A", "t_C"])
    assert d["t_B"] == pytest.approx((1.0 + 2.0 - 0.4) / 2)
    d = derive_times(dict(base, t_A=1.0, t_B=1.2, t_C=2.0, t_out=9.0), ["t_A", "t_B", "t_C"])
    assert d["t_out"] == 9.0
    with pytest.raises(ValueError):
        derive_times(dict(base, t_C=1.0), ["t_C"])


def test_update_best_model_semantics(tmp_path):
    import yaml
    from itrails_b200.yaml_helpers import load_config, update_best_model
    f = tmp_path / "m.best_model.yaml"
    f.write_text(yaml.dump({"fixed_parameters": {"mu": 2e-8}, "optimized_parameters": {},
                            "results": {"log_likelihood": None, "iteration": None}}))
    update_best_model(str(f), ["t_1", "r"], [0.004, 0.5], -100.0, 0)
    d = load_config(str(f))
    assert d["results"] == {"log_likelihood": -100.0, "iteration": 0}
    assert d["optimized_parameters"]["t_1"] == pytest.approx(0.004 / 2e-8)
    assert d["optimized_parameters"]["r"] == pytest.approx(0.5 * 2e-8)
    update_best_model(str(f), ["t_1", "r"], [0.1, 0.1], -200.0, 1)     # worse: unchanged
    assert load_config(str(f))["results"]["iteration"] == 0
    update_best_model(str(f), ["t_1", "r"], [0.1, 0.1], -50.0, 2)
    assert load_config(str(f))["results"]["iteration"] == 2


def test_lpt_partition_balanced_and_deterministic():
    from itrails_b200.distributed import lpt_partition
    rng = np.random.default_rng(1)
    lens = rng.integers(50_000, 150_000, size=100)
    for w in (1, 2, 4, 8):
        parts = lpt_partition(lens, w)
        assert sorted(np.concatenate(parts).tolist()) == list(range(100))
        loads = np.array([lens[p].sum() for p in parts])
        assert loads.max() / loads.mean() < 1.05
        assert all(np.array_equal(x, y) for x, y in zip(parts, lpt_partition(lens, w)))


def test_gloo_world_size_2_allreduce(tmp_path):
    """N>1 host path on CPU: two ranks shard the blocks, compute partials (here with a
    stand-in per-block value) and all-reduce; the sum equals the single-process sum."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        from itrails_b200 import distributed as D
        dist.init_process_group("gloo")
        rng = np.random.default_rng(0)
        V_lst = [rng.integers(0, 625, size=n) for n in rng.integers(10, 500, size=37)]
        local, ids = D.shard_blocks(V_lst)
        assert len(local) == len(ids) and all(local[i] is V_lst[j] for i, j in enumerate(ids))
        partial = np.array([sum(float(v.sum()) for v in local), float(len(local))])
        tot = D.allreduce_sum(partial)
        assert tot[1] == 37 and tot[0] == sum(float(v.sum()) for v in V_lst), tot
        assert D.allreduce_max(dist.get_rank() + 1.5) == 2.5
        print("rank", dist.get_rank(), "ok")
    """))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_workflow_parameter_scaling_and_cases():
    """prepare_optimize on the reference's example config (examples/example_config.yaml
    values): variables, scaled starting values and bounds (workflow_optimize.py:369-405)."""
    from itrails_b200.workflows import prepare_decode, prepare_optimize
    cfg = {"fixed_parameters": {"mu": 1e-8},
           "optimized_parameters": {"N_AB": [50000, 5000, 500000], "N_ABC": [50000, 5000, 500000],
                                    "t_1": [240000, 24000, 2400000], "t_2": [40000, 4000, 400000],
                                    "t_3": [800000, 80000, 8000000],
                                    "t_upper": [745069.3855, 74506.9385, 7450693.8556], "r": [1e-8, 1e-9, 1e-7]},
           "settings": {"n_int_AB": 3, "n_int_ABC": 3}}
    names, start, bounds, fixed, case = prepare_optimize(cfg, 3, 3)
    assert names == ["t_1", "t_2", "N_ABC", "N_AB", "r", "t_upper"] and case == frozenset(["t_1"])
    np.testing.assert_allclose(start, [240000e-8, 40000e-8, 5e-4, 5e-4, 1.0, 745069.3855e-8])
    np.testing.assert_allclose(bounds[4], (0.1, 10.0))
    assert fixed == {"n_int_AB": 3, "n_int_ABC": 3}
    # t_upper derived from t_3 when absent
    cfg2 = {k: dict(v) for k, v in cfg.items()}
    del cfg2["optimized_parameters"]["t_upper"]
    with pytest.raises(ValueError, match="cannot be negative"):      # 80000 - 1.0986 * 500000 < 0
        prepare_optimize(cfg2, 3, 3)
    cfg2["optimized_parameters"]["N_ABC"] = [50000, 40000, 60000]
    cfg2["optimized_parameters"]["t_3"] = [800000, 700000, 900000]
    names2, start2, bounds2, _, _ = prepare_optimize(cfg2, 3, 3)
    last = -np.log1p(-2 / 3)
    assert names2[-1] == "t_upper"
    np.testing.assert_allclose(start2[-1], (800000 - last * 50000) * 1e-8)
    np.testing.assert_allclose(bounds2[-1], ((700000 - last * 60000) * 1e-8, (900000 - last * 40000) * 1e-8))
    bad = {k: dict(v) for k, v in cfg.items()}
    bad["optimized_parameters"]["t_A"] = [1, 1, 1]
    bad["optimized_parameters"]["t_B"] = [1, 1, 1]
    with pytest.raises(ValueError, match="Invalid combination"):
        prepare_optimize(bad, 3, 3)
    # decoding: a best_model-style config (scalars), equals synth.example_model_args
    from itrails_b200 import synth
    dec = {"fixed_parameters": {"mu": 1e-8},
           "optimized_parameters": {"N_AB": 50000, "N_ABC": 50000, "t_1": 240000, "t_2": 40000,
                                    "t_upper": 745069.3855, "r": 1e-8},
           "settings": {"n_int_AB": 3, "n_int_ABC": 3}}
    d, nAB, nABC, aAB, aABC = prepare_decode(dec)
    want = synth.example_model_args(3)
    got = [d[k] for k in ("t_A", "t_B", "t_C", "t_2", "t_upper", "t_out", "N_AB", "N_ABC", "r")]
    np.testing.assert_allclose(got, want, rtol=1e-14)
    assert len(nAB) == 4 and len(nABC) == 4 and nABC[-1] == float("inf") and aABC[0] == 280000.0
    np.testing.assert_allclose(nAB[-1], 40000 / 50000)


def test_viterbi_segments_match_reference_loop():
    """Run-length writer vs a literal restatement of workflow_viterbi.py:698-743."""
    from itrails_b200.workflows import viterbi_segments
    rng = np.random.default_rng(2)

    def ref_loop(res, coords):
        rows = []
        if coords is None:
            seg, cur = 0, res[0]
            for pos in range(1, len(res)):
                if res[pos] != cur:
                    rows.append((seg, pos - 1, cur))
                    seg, cur = pos, res[pos]
            rows.append((seg, len(res) - 1, cur))
            return rows
        first = next((i for i, x in enumerate(coords) if x != -9), None)
        if first is None:
            return rows
        seg = cnn = coords[first]
        cur = res[first]
        for pos in range(first, len(res)):
            if seg == -9:
                seg = coords[pos]; cur = res[pos]; cnn = seg
                continue
            if res[pos] != cur:
                rows.append((seg, cnn, cur))
                seg = coords[pos]; cur = res[pos]
            cnn = coords[pos] if coords[pos] != -9 else cnn
        if not (seg == cnn == -9):
            rows.append((seg, cnn, cur))
        return rows

    for trial in range(30):
        n = int(rng.integers(1, 60))
        res = np.repeat(rng.integers(0, 4, size=n), rng.integers(1, 5, size=n)).astype(np.float64)
        assert viterbi_segments(res) == ref_loop(res, None)
        coords = np.arange(100, 100 + len(res)).tolist()
        for i in rng.integers(0, len(res), size=len(res) // 3):
            coords[int(i)] = -9
        assert viterbi_segments(res, coords) == ref_loop(res, coords)
    assert viterbi_segments(np.array([1.0, 1.0]), [-9, -9]) == []


def test_write_maf_roundtrip(tmp_path):
    import itrails_b200 as itb
    from itrails_b200 import synth
    rng = np.random.default_rng(4)
    V_lst = [rng.integers(0, 625, size=int(T)) for T in (1, 50, 333)]
    p = tmp_path / "rt.maf"
    synth.write_maf(str(p), V_lst)
    back = itb.maf_parser(str(p), ["hg38", "panTro5", "gorGor5", "ponAbe2"])
    assert len(back) == 3 and all(np.array_equal(x, y) for x, y in zip(back, V_lst))
    co = itb.parse_coordinates(str(p), ["hg38", "panTro5", "gorGor5", "ponAbe2"], "hg38")
    assert len(co[1]) == 50


def _torchrun(script, port, nproc=2, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                          capture_output=True, text=True, timeout=timeout, env=e)


def test_gloo_world_size_2_wrappers_shard_and_reduce(tmp_path):
    """The reference-style wrappers under torch.distributed (gloo, 2 ranks, CPU): every
    rank loads only its LPT share of the blocks, the log-likelihood is all-reduced to the
    full sum, the decoders return ONE RESULT PER BLOCK OF V_lst IN INPUT ORDER on every
    rank (the reference's contract, optimizer.py:241-262, 357-377), a rank whose share is
    empty (fewer blocks than ranks) neither crashes nor hangs the all-reduce, and only
    rank 0 writes the optimiser's files.  The GPU engine is replaced by a stand-in that
    answers with the oracle (tests/fake_engine.py), so this covers the host logic of the
    N > 1 path."""
    script = tmp_path / "w2.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "oracle")); sys.path.insert(0, os.path.join({ROOT!r}, "tests"))
        import numpy as np, torch.distributed as dist, yaml
        import hmm_oracle as ho
        from fake_engine import FakeEngine
        from itrails_b200 import distributed as D, engine_cache
        import itrails_b200.optimizer as O
        engine_cache._ENGINE = FakeEngine(int(os.environ["LOCAL_RANK"]))

        dist.init_process_group("gloo")
        rank = dist.get_rank()
        g = np.load(os.path.join({ROOT!r}, "tests", "golden", "model_1_1_example.npz"))
        a, b, pi = g["a"], g["b"], g["pi"]
        rng = np.random.default_rng(1)
        V_lst = [rng.integers(0, 256, size=int(n)) for n in rng.integers(5, 60, size=9)]
        ll = O.loglik_wrapper(a, b, pi, V_lst)
        want = ho.loglik_wrapper(a, b, pi, V_lst)
        assert abs(ll - want) <= 1e-12 * abs(want), (ll, want)
        ll2 = O.loglik_wrapper(a, b, pi, V_lst)            # second call: blocks stay resident
        assert engine_cache._ENGINE.loads == 1 and ll2 == ll
        mine = D.lpt_partition([len(v) for v in V_lst], 2)[rank]
        assert [len(v) for v in engine_cache._ENGINE.V] == [len(V_lst[i]) for i in mine]
        # decoders: the full result, in input order, on both ranks
        paths = O.viterbi_wrapper(a, b, pi, V_lst)
        ref = ho.viterbi_wrapper(a, b, pi, V_lst)
        assert len(paths) == len(V_lst) and all(p.dtype == np.float64 and np.array_equal(p, r) for p, r in zip(paths, ref))
        post = O.post_prob_wrapper(a, b, pi, V_lst)
        refp = ho.post_prob_wrapper(a, b, pi, V_lst)
        assert len(post) == len(V_lst) and all(p.shape == r.shape and np.array_equal(p, r) for p, r in zip(post, refp))
        # fewer blocks than ranks: rank 1's share is empty
        one = [V_lst[3]]
        assert len(D.lpt_partition([len(one[0])], 2)[1]) == 0
        ll1 = O.loglik_wrapper(a, b, pi, one)
        w1 = ho.loglik_wrapper(a, b, pi, one)
        assert abs(ll1 - w1) <= 1e-12 * abs(w1)
        p1 = O.viterbi_wrapper(a, b, pi, one)
        assert len(p1) == 1 and np.array_equal(p1[0], ref[3])
        q1 = O.post_prob_wrapper(a, b, pi, one)
        assert len(q1) == 1 and np.array_equal(q1[0], refp[3])
        # objective: rank 0 alone writes the history / best-model files
        d = dict(zip(("t_A", "t_B", "t_C", "t_2", "t_upper", "t_out", "N_AB", "N_ABC", "r"), g["args"]))
        d.update(n_int_AB=1, n_int_ABC=1)
        res = os.path.join({str(tmp_path)!r}, "run")
        if rank == 0:
            with open(res + ".best_model.yaml", "w") as fh:
                yaml.dump({{"fixed_parameters": {{"mu": 1e-8}}, "optimized_parameters": {{}},
                           "results": {{"log_likelihood": None, "iteration": None}}}}, fh)
        dist.barrier()
        val = O.optimization_wrapper([d["N_AB"]], ["N_AB"], frozenset(["t_A", "t_B", "t_C"]), d, V_lst, res, {{"Nfeval": 0, "time": 0.0}})
        assert abs(-val - want) <= 1e-9 * abs(want), (val, want)
        v1 = O.optimization_wrapper([d["N_AB"]], ["N_AB"], frozenset(["t_A", "t_B", "t_C"]), d, one, res, {{"Nfeval": 1, "time": 0.0}})
        assert abs(-v1 - w1) <= 1e-9 * abs(w1), (v1, w1)
        dist.barrier()
        if rank == 0:
            lines = open(res + ".optimization_history.csv").read().strip().splitlines()
            assert len(lines) == 2 and lines[0].startswith("0,") and lines[1].startswith("1,")
        # batched simplex on the sharded objective: every rank walks the same simplex (the
        # all-reduced values are identical), rank 0 alone writes the files; same evaluations
        # as the sequential search
        from scipy.optimize import minimize
        case = frozenset(["t_A", "t_B", "t_C"])
        bounds = [(d["N_AB"] * 0.2, d["N_AB"] * 5)]
        res2 = os.path.join({str(tmp_path)!r}, "bat")
        if rank == 0:
            with open(res2 + ".best_model.yaml", "w") as fh:
                yaml.dump({{"fixed_parameters": {{"mu": 1e-8}}, "optimized_parameters": {{}},
                           "results": {{"log_likelihood": None, "iteration": None}}}}, fh)
        dist.barrier()
        rb = O.optimizer(["N_AB"], [d["N_AB"] * 1.3], bounds, d, V_lst, res2, case, method="Nelder-Mead-batched")
        assert D.allreduce_max(float(rb.x[0])) == rb.x[0] and D.allreduce_max(-float(rb.x[0])) == -rb.x[0]   # identical on both ranks
        seq = minimize(lambda x: -float(O.loglik_sweep([x], ["N_AB"], case, d, V_lst)[0]), [d["N_AB"] * 1.3],
                       method="Nelder-Mead", bounds=bounds, options={{"maxiter": 10000}})
        assert seq.nfev == rb.nfev and seq.nit == rb.nit and seq.x[0] == rb.x[0], (seq.nfev, rb.nfev)
        assert rb.nbatch < rb.nfev
        dist.barrier()
        if rank == 0:
            lines = open(res2 + ".optimization_history.csv").read().strip().splitlines()
            assert lines[0] == "n_eval,N_AB,loglik,time" and len(lines) == rb.nfev + 1
            best = yaml.safe_load(open(res2 + ".best_model.yaml"))
            assert abs(best["results"]["log_likelihood"] + rb.fun) <= 1e-12 * abs(rb.fun)
        print("rank", rank, "ok")
    """))
    r = _torchrun(script, 29613)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def test_sharded_decode_clis_write_the_single_gpu_files(tmp_path):
    """itrails-viterbi / itrails-posterior with the blocks sharded over GPUs must write the
    files a single GPU writes (workflow_viterbi.py:688-743, workflow_posterior.py:693-716:
    global block indices, reference coordinates of the right block, one writer): run the
    CLIs (a) in one process on one engine, (b) under torchrun with 2 gloo ranks, (c) in one
    process driving two engines (``--n_gpu 2``), and compare the CSVs byte for byte — the
    posterior both through the native part-file writer and through the csv.writer loop."""
    from itrails_b200 import synth
    m = golden("model_2_2_example.npz")
    rng = np.random.default_rng(12)
    V_lst = [ho.sample_block(m["a"], m["b"], m["pi"], int(T), rng, p_n=0.05) for T in (3000, 410, 1, 2220, 970, 1600, 350)]
    species = ["hg38", "panTro5", "gorGor5", "ponAbe2"]
    maf = tmp_path / "s.maf"
    synth.write_maf(str(maf), V_lst, species)
    script = tmp_path / "cli.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "oracle")); sys.path.insert(0, os.path.join({ROOT!r}, "tests"))
        from fake_engine import FakeEngine
        from itrails_b200 import engine_cache, workflows
        import itrails_b200.engine_cache as ec
        ec.Engine = FakeEngine                      # every engine the package creates is the stand-in
        mode, out = sys.argv[1], sys.argv[2]
        extra = []
        if mode == "torchrun":
            import torch.distributed as dist
            dist.init_process_group("gloo")
        elif mode == "ngpu2":
            extra = ["--n_gpu", "2"]
        common = ["--input", {str(maf)!r}, "--mu", "1e-8", "--t1", "240000", "--t2", "40000", "--t_upper", "745069.3855",
                  "--N_AB", "50000", "--N_ABC", "50000", "--r", "1e-8", "--n_int_AB", "2", "--n_int_ABC", "2",
                  "--species_list", *{species!r}, "--reference", "hg38"] + extra
        workflows.viterbi_main(common + ["--output", os.path.join(out, "v")])
        workflows.posterior_main(common + ["--output", os.path.join(out, "p")])
        os.environ["ITRAILS_PY_CSV"] = "1"
        workflows.posterior_main(common + ["--output", os.path.join(out, "q")])
        print("done")
    """))
    outs = {}
    for mode in ("single", "torchrun", "ngpu2"):
        out = tmp_path / mode
        out.mkdir()
        if mode == "torchrun":
            r = _torchrun(f"{script} torchrun {out}".split()[0], 29617) if False else subprocess.run(
                [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                 "--master-port", "29617", str(script), "torchrun", str(out)], capture_output=True, text=True, timeout=600)
        else:
            env = {k: v for k, v in os.environ.items() if k not in ("LOCAL_RANK", "RANK", "WORLD_SIZE", "ITRAILS_DEVICE")}
            r = subprocess.run([sys.executable, str(script), mode, str(out)], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, mode + r.stdout[-3000:] + r.stderr[-3000:]
        outs[mode] = {f: open(out / f, "rb").read() for f in ("v.viterbi.csv", "p.posterior.csv", "q.posterior.csv")}
        assert not [f for f in os.listdir(out) if ".part" in f], "part files must be removed after the splice"
    single = outs["single"]
    assert single["p.posterior.csv"] == single["q.posterior.csv"]
    assert single["v.viterbi.csv"].count(b"\n") >= 8 and single["p.posterior.csv"].count(b"\n") == 1 + sum(len(v) for v in V_lst)
    for mode in ("torchrun", "ngpu2"):
        for f, data in outs[mode].items():
            assert data == single[f], f"{mode}: {f} differs from the single-GPU file"


def test_native_posterior_csv_is_byte_identical_to_csv_writer(tmp_path):
    """itr_csv_posterior_host / itr_csv_format_double (host C++, no GPU) against Python's
    csv.writer + repr(float), which is what workflow_posterior.py:697-716 produces."""
    import csv
    import ctypes
    import io
    from itrails_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(64)
    rng = np.random.default_rng(5)
    vals = [0.0, -0.0, 1.0, 0.1, 1e-4, 1e-5, 9.999e-5, 1e15, 1e16, 1.5e16, 9999999999999998.0, 1e22, 1e23, 5e-324,
            2.2250738585072014e-308, 1.7976931348623157e308, float("inf"), -float("inf"), 1 / 3, 1e-7, 1e100, 1e-100]
    vals += list(rng.random(20000)) + list(np.exp(rng.uniform(-700, 700, 20000)))
    vals += [v for v in np.frombuffer(rng.bytes(8 * 20000), dtype=np.float64) if v == v]
    for v in vals:
        n = lib.itr_csv_format_double(float(v), buf, 64)
        assert buf.value.decode() == repr(float(v)) and n == len(repr(float(v)))
    assert lib.itr_csv_format_double(1.0, buf, 8) < 0
    K, lens = 27, [5, 1, 4097, 9000]
    off = np.zeros(len(lens) + 1, np.int64)
    off[1:] = np.cumsum(lens)
    post = rng.random((off[-1], K)) ** 8
    post /= post.sum(1, keepdims=True)
    post[3, :] = 0
    post[3, 2] = 1.0
    post[10, 5] = 1e-300
    pos = rng.integers(-9, 10**9, off[-1]).astype(np.int64)
    for positions in (None, pos):
        p = str(tmp_path / "x.csv")
        rc = lib.itr_csv_posterior_host(p.encode(), K, len(lens), _lib.as_ptr(off, ctypes.c_int64),
                                        _lib.as_ptr(positions, ctypes.c_int64), _lib.as_ptr(post, ctypes.c_double), 3)
        assert rc == 0
        s = io.StringIO(newline="")
        w = csv.writer(s)
        w.writerow(["alignment_block_idx", "position_idx"] + [f"prob_state_{i}" for i in range(K)])
        for bi in range(len(lens)):
            for r in range(off[bi], off[bi + 1]):
                w.writerow([bi, int(pos[r]) if positions is not None else r - off[bi]] + post[r].tolist())
        assert open(p, "rb").read() == s.getvalue().encode()
    assert lib.itr_csv_posterior_host(str(tmp_path / "no" / "dir.csv").encode(), K, 0, None, None, None, 1) == _lib.ITR_ERR_IO


# ---------------------------------------------------------------------------
# Speculative (batched) Nelder-Mead against the sequential search the reference runs
# (scipy.optimize.minimize(method="Nelder-Mead", bounds=...), optimizer.py:623-637)
# ---------------------------------------------------------------------------
def _nm_objectives():
    def rosen(x):
        return float(np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))

    def crinkled(x):                      # non-smooth: forces contractions and shrinks
        return float(np.sum(np.abs(x - 0.3)) + 0.5 * np.abs(np.sin(7 * x)).sum() + np.max(np.abs(x)))

    def zero_start(x):
        return float(np.sum((x - np.array([0.5, -0.25, 2.0])) ** 2))
    return [
        (rosen, np.array([-1.2, 1.0, 0.8, 1.9, -0.5]), [(-2.0, 2.0)] * 5, {}),
        (rosen, np.array([1.99, 1.0, 0.8]), [(-2.0, 2.0)] * 3, {}),          # simplex reflected at the upper bound
        (crinkled, np.array([1.0, -1.0, 0.5, 0.25]), [(-1.5, 1.5)] * 4, {}),
        (crinkled, np.array([1.0, -1.0, 0.5, 0.25]), None, {"maxiter": 40}),
        (crinkled, np.array([1.0, -1.0, 0.5, 0.25]), [(-1.5, 1.5)] * 4, {"maxfev": 37}),
        (crinkled, np.array([1.0, -1.0]), [(-1.5, 1.5)] * 2, {"maxfev": 2}),
        (zero_start, np.array([0.0, 0.0, 1.0]), [(-3.0, 3.0)] * 3, {"maxiter": 10000}),
    ]


@pytest.mark.parametrize("case", range(7))
def test_batched_nelder_mead_reproduces_sequential_search(case):
    import warnings
    from scipy.optimize import minimize
    from itrails_b200.batched_simplex import minimize_neldermead_batched
    fun, x0, bounds, opts = _nm_objectives()[case]
    seq = []

    def logged(x):
        f = fun(x)
        seq.append((np.array(x, copy=True), f))
        return f
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = minimize(logged, x0, method="Nelder-Mead", bounds=bounds, options=dict(opts))
    got = []
    res = minimize_neldermead_batched(lambda X: [fun(x) for x in X], x0, bounds=bounds,
                                      consume=lambda x, f: got.append((np.array(x, copy=True), f)), **opts)
    assert res.nfev == ref.nfev == len(got) == len(seq)
    assert res.nit == ref.nit and res.status == ref.status
    for (xa, fa), (xb, fb) in zip(got, seq):
        assert np.array_equal(xa, xb) and fa == fb
    assert np.array_equal(res.x, ref.x) and res.fun == ref.fun
    assert np.array_equal(res.final_simplex[0], ref.final_simplex[0])
    if ref.nit > 5:
        assert res.nbatch < ref.nfev            # fewer dependent objective calls than evaluations
