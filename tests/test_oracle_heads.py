"""Pins oracle/heads.py and oracle/metrics.py against fixtures generated from the
reference's own losses.py / metrics.py (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import heads as oh
from oracle import metrics as om

CASES = {
    "focal_g1": (oh.LOSS_FOCAL, 1.0, False),
    "focal_g2": (oh.LOSS_FOCAL, 2.0, False),
    "focal_g0p5_alpha": (oh.LOSS_FOCAL, 0.5, True),
    "focal_g2_alpha": (oh.LOSS_FOCAL, 2.0, True),
    "ce": (oh.LOSS_CE, 0.0, False),
    "ce_weight": (oh.LOSS_CE, 0.0, True),
}


def load_heads(golden_dir):
    g = np.load(golden_dir / "heads_golden.npz")
    T = 3
    emb = torch.from_numpy(g["emb"])
    Ws = [torch.from_numpy(g[f"W{t}"]) for t in range(T)]
    bs = [torch.from_numpy(g[f"b{t}"]) for t in range(T)]
    alphas = [torch.from_numpy(g[f"alpha{t}"]) for t in range(T)]
    labels = torch.from_numpy(g["labels"])
    return g, emb, Ws, bs, alphas, labels


@pytest.mark.parametrize("case", sorted(CASES))
def test_heads_oracle_matches_reference_f64(golden_dir, case):
    g, emb, Ws, bs, alphas, labels = load_heads(golden_dir)
    kind, gamma, weighted = CASES[case]
    r = oh.heads_loss_fwd_bwd(emb, Ws, bs, labels, kind, gamma, alphas if weighted else None, dtype=torch.float64)
    exp_loss = g[f"{case}.f64.loss"]
    got = np.array([float(x) for x in r["loss"]] + [float(r["total"])])
    np.testing.assert_allclose(got, exp_loss, rtol=1e-12, atol=1e-14)
    for t in range(3):
        np.testing.assert_allclose(r["dW"][t].numpy(), g[f"{case}.f64.dW{t}"], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(r["db"][t].numpy(), g[f"{case}.f64.db{t}"], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(r["logits"][t].numpy(), g[f"{case}.f64.logits{t}"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(r["demb"].numpy(), g[f"{case}.f64.demb"], rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("case", sorted(CASES))
def test_unnormalised_sums_reassemble_the_mean(golden_dir, case):
    """sum / denom over row shards == the reference's global mean: the K4 contract."""
    g, emb, Ws, bs, alphas, labels = load_heads(golden_dir)
    kind, gamma, weighted = CASES[case]
    cw = alphas if weighted else None
    parts = [oh.unnormalised_sums(emb[a:b], Ws, bs, labels[a:b], kind, gamma, cw) for a, b in ((0, 5), (5, 20), (20, 37))]
    for t in range(3):
        den = sum(p["denom"][t] for p in parts)
        loss = sum(p["loss_sum"][t] for p in parts) / den
        dW = sum(p["dW_sum"][t] for p in parts) / den
        np.testing.assert_allclose(float(loss), g[f"{case}.f64.loss"][t], rtol=1e-12)
        np.testing.assert_allclose(dW.numpy(), g[f"{case}.f64.dW{t}"], rtol=1e-9, atol=1e-14)


def test_all_ignored_task_is_zero(golden_dir):
    g = np.load(golden_dir / "heads_golden.npz")
    z = torch.randn(5, 3)
    assert float(oh.focal_loss(z, torch.full((5,), -100), None, 1.0)) == float(g["all_ignored.loss"]) == 0.0


def test_metrics_oracle_matches_reference(golden_dir):
    g = np.load(golden_dir / "metrics_golden.npz")
    names = ["a_color", "b_size", "c_kind"]
    res = {"running_loss": {}, "confidences": {}, "predictions": {}, "ground_truth": {}}
    for n in names:
        z = g[f"{n}.logits"]
        res["confidences"][n] = torch.from_numpy(z).softmax(-1, dtype=torch.float32).numpy().tolist()
        pred = om.argmax_first_fast(z)
        assert np.array_equal(pred, om.argmax_first(z))
        res["predictions"][n] = pred.tolist()
        res["ground_truth"][n] = g[f"{n}.gt"].tolist()
        res["running_loss"][n] = g[f"{n}.running_loss"].tolist()
        # balanced accuracy from the integer confusion matrix == sklearn, bit for bit
        cm = om.confusion_matrix(g[f"{n}.gt"], pred, z.shape[1])
        assert om.balanced_accuracy_from_cm(cm) == float(g[f"{n}.epoch_acc"])
    res["running_loss"]["loss"] = g["loss"].tolist()
    m = om.compute_metrics("multi", res, names)
    assert m["epoch_acc"] == float(g["epoch_acc"])
    for n in names:
        assert m[n]["epoch_acc"] == float(g[f"{n}.epoch_acc"])
        np.testing.assert_array_equal(np.asarray(m[n]["epoch_roc_auc"], dtype=np.float64), g[f"{n}.epoch_roc_auc"])
        assert m[n]["epoch_loss"] == float(g[f"{n}.epoch_loss"])


def test_argmax_ties_and_nan():
    z = np.array([[1.0, 3.0, 3.0], [np.nan, 5.0, np.nan], [2.0, np.nan, 9.0], [-np.inf, -np.inf, -np.inf]], np.float32)
    exp = torch.from_numpy(z).argmax(-1).numpy()
    assert np.array_equal(om.argmax_first(z), exp)
    assert list(exp) == [1, 0, 1, 0]


def test_oracle_roc_auc_counts_equal_sklearn():
    """The integer Mann-Whitney form the device kernel (K5) is checked against == sklearn.roc_auc_score, including
    heavy ties (quantised scores) and the reference's one-vs-rest use (metrics.py:33-42)."""
    from sklearn.metrics import roc_auc_score
    from sklearn.preprocessing import label_binarize
    from oracle import metrics as om
    from nkb_classification_b200.metrics import roc_auc_from_counts
    rng = np.random.default_rng(3)
    for N, classes, quant in [(500, (3, 2), None), (2000, (7,), 16), (64, (2,), 4), (300, (4, 4), 2)]:
        seg = np.concatenate([[0], np.cumsum(classes)]).tolist()
        z = rng.normal(size=(N, seg[-1])).astype(np.float32)
        probs = np.concatenate([np.exp(z[:, a:b]) / np.exp(z[:, a:b]).sum(1, keepdims=True)
                                for a, b in zip(seg[:-1], seg[1:])], 1).astype(np.float32)
        if quant:
            probs = (np.round(probs * quant) / quant).astype(np.float32)
        labels = np.stack([rng.integers(0, c, N) for c in classes], 1)
        counts = om.roc_auc_counts(probs, labels, seg)
        for t, C in enumerate(classes):
            gt, conf = labels[:, t], probs[:, seg[t]:seg[t + 1]]
            got = roc_auc_from_counts(counts[seg[t]:seg[t + 1]], C)
            if C > 2:
                gb = label_binarize(gt, classes=range(C))
                exp = np.array([roc_auc_score(gb[:, c], conf[:, c]) for c in range(C)])
            else:
                exp = roc_auc_score(gt, conf[:, 1])
            assert np.allclose(got, exp, rtol=0, atol=1e-12), (N, classes, quant, t)
    # a class absent from the ground truth -> NaN there; a single class present -> all NaN (metrics.py:34-35)
    c = np.array([[10, 5, 5], [0, 0, 10], [8, 5, 5]])
    r = roc_auc_from_counts(c, 3)
    assert np.isnan(r[1]) and r[0] == 10 / 50 and r[2] == 8 / 50
    assert np.isnan(roc_auc_from_counts(np.array([[0, 9, 0], [0, 0, 9], [0, 0, 9]]), 3)).all()
    assert np.isnan(roc_auc_from_counts(np.array([[0, 0, 9], [0, 9, 0]]), 2))
