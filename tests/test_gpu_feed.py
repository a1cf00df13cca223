"""GPU parity tests of the device-side training feed against oracle/feed.py:
gathers bit-exact, negatives obey the reference sampler's contract and distribution."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _toy(n_users=200, n_items=500, n_inter=20_000, seed=0):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, n_users - 5, n_inter)            # the last users have no interactions at all
    i = np.minimum(rng.zipf(1.3, n_inter) - 1, n_items - 1)
    lab = rng.integers(0, 2, n_inter).astype(np.float64)
    uf = rng.standard_normal((n_users, 3)).astype(np.float32)
    mf = rng.standard_normal((n_items, 20)).astype(np.float32)
    pos = {}
    for a, b in zip(u.tolist(), i.tolist()):
        pos.setdefault(a, []).append(b)
    return u, i, lab, uf, mf, pos


def test_feed_batches_match_the_oracle():
    from b200rec.feed import DeviceInteractionFeed
    from oracle.feed import make_batch
    u, i, lab, uf, mf, pos = _toy()
    R, B = 16, 1024
    feed = DeviceInteractionFeed(u, i, lab, uf, mf, pos, num_items=mf.shape[0], num_negatives=R, batch_size=B,
                                 shuffle=False, seed=7)
    assert len(feed) == (len(u) + B - 1) // B
    seen = 0
    for bi, batch in enumerate(feed):
        rows = list(range(bi * B, min(len(u), (bi + 1) * B)))
        ref = make_batch(rows, u, i, lab, uf, mf, pos, mf.shape[0], R, True)
        got = {k: v.cpu().numpy() for k, v in batch.items()}
        assert list(got) == list(ref)                                    # same keys in collate_fn's order
        for k in ("user_idx", "pos_item_idx", "label", "user_features", "pos_item_features"):
            assert got[k].dtype == ref[k].dtype and np.array_equal(got[k], ref[k]), k   # gathers are bit exact
        neg = got["neg_item_indices"]
        assert neg.shape == ref["neg_item_indices"].shape and neg.dtype == np.int64
        assert np.array_equal(got["neg_item_features"], mf[neg])        # features of the sampled items, bit exact
        for uu, row in zip(got["user_idx"].tolist(), neg.tolist()):     # the reference sampler's contract
            assert len(set(row)) == R and not (set(row) & set(pos.get(uu, []))) and min(row) >= 0 and max(row) < mf.shape[0]
        seen += len(rows)
    assert seen == len(u)
    feed.check()
    assert got["user_idx"].shape[0] == len(u) % B                        # ragged last batch, like the DataLoader


def test_feed_shuffle_covers_every_interaction_once_and_is_seeded():
    from b200rec.feed import DeviceInteractionFeed
    u, i, lab, uf, mf, pos = _toy(seed=3)
    mk = lambda s: DeviceInteractionFeed(u, i, lab, uf, mf, pos, num_items=mf.shape[0], num_negatives=4, batch_size=777, seed=s)
    a, b, c = mk(5), mk(5), mk(6)
    ea = [{k: v.cpu().numpy() for k, v in bt.items()} for bt in a]
    eb = [{k: v.cpu().numpy() for k, v in bt.items()} for bt in b]
    ec = [{k: v.cpu().numpy() for k, v in bt.items()} for bt in c]
    pairs = np.concatenate([np.stack([x["user_idx"], x["pos_item_idx"]], 1) for x in ea])
    want = np.stack([u, i], 1)
    assert np.array_equal(pairs[np.lexsort(pairs.T)], want[np.lexsort(want.T)])      # a permutation of the table
    assert not np.array_equal(pairs, want)                                            # ... that is shuffled
    assert all(np.array_equal(x[k], y[k]) for x, y in zip(ea, eb) for k in x)         # same seed, same epoch
    assert not all(np.array_equal(x["neg_item_indices"], y["neg_item_indices"]) for x, y in zip(ea, ec))
    e2 = [bt["neg_item_indices"].cpu().numpy() for bt in a]                           # second epoch: new order and draws
    assert not np.array_equal(e2[0], ea[0]["neg_item_indices"])


def test_negative_sampler_is_uniform_over_the_pool():
    """One user with 40 positives out of 100 items, 200k rows x 4 draws: every admissible item is hit equally often
    (chi-square against the uniform law over the 60-item pool, 59 degrees of freedom)."""
    from b200rec import kernels as K
    n_items, R, B = 100, 4, 200_000
    positives = np.sort(np.random.default_rng(1).choice(n_items, 40, replace=False)).astype(np.int32)
    indptr = torch.tensor([0, len(positives)], dtype=torch.int64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    neg = K.sample_negatives(torch.zeros(B, dtype=torch.int64, device="cuda"), indptr, torch.from_numpy(positives).cuda(),
                             n_items, R, 99, 0, err).cpu().numpy()
    assert int(err.item()) == 0
    counts = np.bincount(neg.reshape(-1), minlength=n_items)
    assert counts[positives].sum() == 0
    pool = np.setdiff1d(np.arange(n_items), positives)
    exp = B * R / len(pool)
    chi2 = float(((counts[pool] - exp) ** 2 / exp).sum())
    assert chi2 < 110.0, chi2          # P(chi2_59 > 110) ~ 6e-5
    assert (np.sort(neg, 1)[:, 1:] != np.sort(neg, 1)[:, :-1]).all()    # without replacement


def test_feed_errors_mirror_the_reference():
    from b200rec import kernels as K
    from b200rec.feed import DeviceInteractionFeed
    u, i, lab, uf, mf, pos = _toy(n_users=10, n_items=12, n_inter=400, seed=9)
    with pytest.raises(ValueError):      # some user has interacted with (almost) everything: pool < num_negatives
        DeviceInteractionFeed(u, i, lab, uf, mf, pos, num_items=12, num_negatives=10)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = K.gather_rows(torch.from_numpy(mf).cuda(), torch.tensor([0, 12, 3], device="cuda"), err)
    assert int(err.item()) == 1 and torch.equal(out[1], torch.zeros(20, device="cuda"))


def test_trainer_epoch_runs_on_the_device_feed():
    """TwoTowerTrainer.train_epoch / validate consume the feed exactly like the reference DataLoader."""
    from b200rec.feed import DeviceInteractionFeed
    from b200rec.trainer import TwoTowerTrainer
    from b200rec.training_utils import create_two_tower_model_for_training
    u, i, lab, uf, mf, pos = _toy(n_inter=6000, seed=11)
    torch.manual_seed(0)
    model = create_two_tower_model_for_training(3, 20, {"embedding_dim": 32, "hidden_layers": [64, 32]})
    train = DeviceInteractionFeed(u, i, lab, uf, mf, pos, num_items=mf.shape[0], num_negatives=4, batch_size=512, seed=1)
    val = DeviceInteractionFeed(u[:1500], i[:1500], lab[:1500], uf, mf, batch_size=512, shuffle=False, is_training=False)
    tr = TwoTowerTrainer(model, train, val, {"checkpoint_dir": "/tmp/b200rec_feed_test"}, device="cuda")
    l1 = tr.train_epoch(1)
    l2 = tr.train_epoch(2)
    v = tr.validate()
    assert np.isfinite([l1, l2, v]).all() and l2 < l1
    train.check()
