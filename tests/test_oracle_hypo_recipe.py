"""Known-answer tests of the oracle's _prepareHypoRecipe! restatement, ported from the
reference's own structural tests: /root/reference/test/testExplicitMultihypo.jl (all testsets).
These pin SURVEY.md §8 row a6 (hypothesis labels / recipe structure)."""
import numpy as np

import oracle as O


def _jl_same(gt, got):
    """The reference compares with `sum(gt .- got) == 0`; Julia broadcasting of an empty vector
    against a 1-element vector is empty, so `Int[]` "equals" `[1]` there.  The code's actual answer
    for a certain sfidx == pidx is sort(union(certainidx, sfidx)) (EDM.jl:205-207, `&&` binds tighter
    than `||`), which is what the oracle returns."""
    if len(gt) == 0 and len(got) == 1:
        return True
    return list(gt) == list(got)


def _partition_ok(r, n):
    for (hyp, _), el in zip(r["activehypo"], r["allelements"]):
        assert el == [i + 1 for i in np.nonzero(r["mhidx"] == hyp)[0]] or el == []
    assert len(r["mhidx"]) == n


def test_only_nullhypothesis():  # testExplicitMultihypo.jl:8-59
    for sf in (1, 2):
        r = O.hypo_recipe(None, 20, sf, 2, [1, 1], 0.5, u=np.random.default_rng(sf).random(20))
        assert r["certainidx"] == [1, 2]
        assert len(r["allelements"][0]) > 3 and len(r["allelements"][1]) > 3
        assert len(r["allelements"][2]) == 0
        assert len(r["allelements"][0]) + len(r["allelements"][1]) == 20
        assert r["activehypo"] == [(0, [sf]), (1, [1, 2]), (2, [])]
        assert (r["mhidx"] == 0).sum() > 3 and (r["mhidx"] == 1).sum() > 3
        assert r["allelements"][0] == [i + 1 for i in np.nonzero(r["mhidx"] == 0)[0]]
        assert r["allelements"][1] == [i + 1 for i in np.nonzero(r["mhidx"] == 1)[0]]


def test_without_multihypothesis():  # :63-115
    for sf in (1, 2):
        r = O.hypo_recipe(None, 20, sf, 2)
        assert r["certainidx"] == [1, 2]
        assert r["allelements"] == [[], list(range(1, 21)), []]
        assert r["activehypo"] == [(0, [sf]), (1, [1, 2]), (2, [])]
        assert np.all(r["mhidx"] == 1)


def test_bimodal_certain_variable():  # :119-160
    r = O.hypo_recipe([0.0, 0.5, 0.5], 40, 1, 3, u=np.random.default_rng(3).random(40))
    assert r["certainidx"] == [1]
    assert r["allelements"][0] == []
    assert len(r["allelements"][1]) > 3 and len(r["allelements"][2]) > 3
    assert len(r["allelements"][1]) + len(r["allelements"][2]) == 40
    assert [h for h, _ in r["activehypo"]] == [1, 2, 3]
    assert _jl_same([], r["activehypo"][0][1]) and r["activehypo"][0][1] == [1]
    assert [v for _, v in r["activehypo"][1:]] == [[1, 2], [1, 3]]
    assert (r["mhidx"] == 2).sum() > 3 and (r["mhidx"] == 3).sum() > 3
    _partition_ok(r, 40)


def test_bimodal_fractional_variable_1_of_2():  # :163-203
    r = O.hypo_recipe([0.0, 0.5, 0.5], 40, 2, 3, u=np.random.default_rng(4).random(40))
    assert r["certainidx"] == [1]
    assert r["activehypo"] == [(0, [2]), (1, [1, 2]), (2, [1, 2]), (3, [2, 3])]
    assert len(r["allelements"][0]) > 1.5 and len(r["allelements"][1]) == 0
    assert len(r["allelements"][2]) > 3 and len(r["allelements"][3]) > 3
    assert sum(len(e) for e in r["allelements"]) == 40
    _partition_ok(r, 40)


def test_bimodal_fractional_variable_2_of_2():  # :205-245
    r = O.hypo_recipe([0.0, 0.5, 0.5], 40, 3, 3, u=np.random.default_rng(5).random(40))
    assert r["certainidx"] == [1]
    assert r["activehypo"] == [(0, [3]), (1, [1, 3]), (2, [2, 3]), (3, [1, 3])]
    assert len(r["allelements"][1]) == 0
    assert sum(len(e) for e in r["allelements"]) == 40
    _partition_ok(r, 40)


def test_trimodal_all_sfidx():  # :290-431
    p = [0.0, 0.33, 0.33, 0.34]
    r = O.hypo_recipe(p, 50, 1, 4, u=np.random.default_rng(6).random(50))
    assert r["certainidx"] == [1]
    assert [h for h, _ in r["activehypo"]] == [1, 2, 3, 4]
    assert _jl_same([], r["activehypo"][0][1])
    assert [v for _, v in r["activehypo"][1:]] == [[1, 2], [1, 3], [1, 4]]
    assert all(len(r["allelements"][k]) > 3 for k in (1, 2, 3)) and r["allelements"][0] == []
    assert sum(len(e) for e in r["allelements"]) == 50
    expect = {
        2: [(0, [2]), (1, [1, 2]), (2, [1, 2]), (3, [2, 3, 4]), (4, [2, 3, 4])],
        3: [(0, [3]), (1, [1, 3]), (2, [2, 3, 4]), (3, [1, 3]), (4, [2, 3, 4])],
        4: [(0, [4]), (1, [1, 4]), (2, [2, 3, 4]), (3, [2, 3, 4]), (4, [1, 4])],
    }
    for sf, ah in expect.items():
        r = O.hypo_recipe(p, 70, sf, 4, u=np.random.default_rng(10 + sf).random(70))
        assert r["certainidx"] == [1]
        assert r["activehypo"] == ah
        assert len(r["allelements"][1]) == 0
        assert all(len(r["allelements"][k]) > 3 for k in (0, 2, 3, 4))
        assert sum(len(e) for e in r["allelements"]) == 70
        _partition_ok(r, 70)


def test_bad_init_null_class_weight():
    """sfidx fractional => class 0 is prepended with weight 1/(U+1)  (EDM.jl:176-183)."""
    u = (np.arange(30000) + 0.5) / 30000
    r = O.hypo_recipe([0.0, 0.5, 0.5], 30000, 2, 3, u=u)
    frac = [(r["mhidx"] == k).mean() for k in range(4)]
    assert np.allclose(frac, [1 / 3, 0.0, 1 / 3, 1 / 3], atol=1e-3)


def test_uninitialised_hypotheses_are_suppressed():
    """fewer than lenXi-1 initialised => uninitialised (non-sf) hypotheses get p=0  (EDM.jl:161-172)."""
    u = np.random.default_rng(1).random(500)
    r = O.hypo_recipe([0.0, 0.25, 0.25, 0.25, 0.25], 500, 1, 5, isinit=[0, 1, 0, 0, 1], u=u)
    assert set(np.unique(r["mhidx"])) == {2, 5}


def test_explicit_labels_pass_through_bit_exact():
    lab = np.random.default_rng(2).integers(2, 4, 64).astype(np.int32)
    r = O.hypo_recipe([0.0, 0.5, 0.5], 64, 1, 3, mhidx_in=lab)
    assert np.array_equal(r["mhidx"], lab)
