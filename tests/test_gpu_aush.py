"""AUSH generator / discriminator step on the CUDA path (recad_b200/attacker.py, csrc/aush.cu) against the golden run of the
live reference (tests/golden/make_golden_aush.py: the unmodified recad.model.attacker.Aush on CPU) and against the
oracle's restatement (oracle/aush.py).  Bars: the numpy generator is left in the reference's state (bit-exact), epoch
losses 1e-4 relative, discriminator parameters 1e-4 of each tensor's largest magnitude after every epoch, the generator
untouched, the generated fake profiles equal (a rounding of a generator output within 1e-6 of x.5 may differ)."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


class StubExplicit:
    """What Aush touches of the attack dataset, restated from recad/dataset/explicit.py:151-188: `info_describe()` with
    n_items / train_mat and the train-mode batch generator (filter -> np.random.permutation -> slices of batch_size)."""

    def __init__(self, train_mat, batch_size, device):
        self.train_mat, self.batch_size, self.device = train_mat, batch_size, device
        self.n_users, self.n_items = train_mat.shape

    def info_describe(self):
        return {"n_users": self.n_users, "n_items": self.n_items, "train_mat": self.train_mat}

    def generate_batch(self, **config):
        user_filter = config.get("user_filter", None)
        idx = user_filter(train_mat=self.train_mat) if user_filter is not None else list(range(len(self.train_mat)))
        idx = np.random.permutation(idx)
        for b in range((len(idx) + self.batch_size - 1) // self.batch_size):
            rows = idx[b * self.batch_size:(b + 1) * self.batch_size]
            yield {"users": torch.tensor(rows, dtype=torch.int64).to(self.device),
                   "users_mat": torch.tensor(self.train_mat[rows, :].astype("float"), dtype=torch.float32).to(self.device)}


def _attacker(case):
    from recad_b200 import model
    dev = torch.device("cuda:0")
    mat, G, D, kw, batch, targets, z = util.aush_case(case)
    att = model.from_config("attacker", "aush", device=dev, **kw).I(dataset=StubExplicit(mat, batch, dev))
    att.load_netG_state(G)
    att.load_netD_state(D)
    return att, targets, z


@pytest.mark.parametrize("case", ["a", "b"])
def test_aush_train_step_and_generate_fake_match_the_reference(case):
    att, targets, z = _attacker(case)
    np.random.set_state(("MT19937", z[f"{case}_np_key_start"], int(z[f"{case}_np_pos_start"]), 0, 0.0))
    gold = z[f"{case}_losses"]
    for e in range(len(gold)):
        loss = att.train_step(target_id_list=targets, input_describe={}, progress_bar=None)
        assert len(loss) == len(att.output_describe()["train_step"])
        assert np.allclose(loss, gold[e], rtol=1e-4, atol=0), (e, loss, gold[e])
        sd = att.netD_state()
        for k, v in sd.items():
            w = z[f"{case}_D{e + 1}__{k}"]
            assert v.shape == w.shape and np.abs(v.cpu().numpy() - w).max() <= 1e-4 * np.abs(w).max(), (e, k)
    for k, v in att.netG_state().items():
        assert np.array_equal(v.cpu().numpy(), z[f"{case}_G1__{k}"]), k
    st = np.random.get_state()
    assert np.array_equal(st[1], z[f"{case}_np_key_mid"]) and st[2] == int(z[f"{case}_np_pos_mid"])
    fake = att.generate_fake(target_id_list=targets)
    st = np.random.get_state()
    assert np.array_equal(st[1], z[f"{case}_np_key_end"]) and st[2] == int(z[f"{case}_np_pos_end"])
    g = z[f"{case}_fake"]
    assert fake.shape == g.shape and fake.dtype == g.dtype
    assert np.mean(fake != g) <= 2.0 / g.size and np.abs(fake - g).max() <= 1.0


def test_aush_step_is_bit_stable_and_matches_the_oracle_on_other_shapes():
    """Another shape (odd filler count, 5 selected items, batch not a multiple of 8, 2 batches) against the oracle, twice:
    the two runs must agree bit for bit (no atomics anywhere in the step)."""
    from oracle import aush as oa
    from recad_b200 import model
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(4)
    U, I = 230, 517
    mat = ((rng.random((U, I)) < 0.12) * rng.integers(1, 6, (U, I))).astype(np.float32)
    kw = dict(selected_ids=[9, 400, 2, 77, 301], filler_num=7, attack_num=11, ZR_ratio=0.35)
    targets = [13]
    runs = []
    for rep in range(2):
        torch.manual_seed(3)
        att = model.from_config("attacker", "aush", device=dev, lr_d=0.002, **kw).I(dataset=StubExplicit(mat, 123, dev))
        G0, D0 = {k: v.cpu().numpy() for k, v in att.netG_state().items()}, {k: v.cpu().numpy() for k, v in att.netD_state().items()}
        np.random.seed(8)
        losses = [att.train_step(target_id_list=targets) for _ in range(3)]
        fake = att.generate_fake(target_id_list=targets)
        runs.append((losses, {k: v.cpu().numpy() for k, v in att.netD_state().items()}, fake))
    assert runs[0][0] == runs[1][0] and np.array_equal(runs[0][2], runs[1][2])
    assert all(np.array_equal(runs[0][1][k], runs[1][1][k]) for k in runs[0][1])
    o = oa.AushOracle(mat, G0, D0, batch_size=123, lr_d=0.002, **kw)
    np.random.seed(8)
    for e in range(3):
        assert np.allclose(runs[0][0][e], o.train_step(targets), rtol=1e-4, atol=0)
    for l in range(4):
        w = o.D.W[l]
        assert np.abs(runs[0][1][f"main.{2 * l}.weight"] - w).max() <= 1e-4 * np.abs(w).max()
        assert np.abs(runs[0][1][f"main.{2 * l}.bias"] - o.D.b[l]).max() <= 1e-4 * np.abs(o.D.b[l]).max()
    ofake = o.generate_fake(targets)
    assert np.mean(runs[0][2] != ofake) <= 2.0 / ofake.size


def test_aush_construction_draws_the_reference_init_stream():
    """aush.py:26-36: generator then discriminator from the global torch CPU generator."""
    from recad_b200 import model
    dev = torch.device("cuda:0")
    mat, _, _, kw, batch, _, _ = util.aush_case("a")
    torch.manual_seed(2023)
    att = model.from_config("attacker", "aush", device=dev, **kw).I(dataset=StubExplicit(mat, batch, dev))
    torch.manual_seed(2023)
    nn, I = torch.nn, mat.shape[1]
    G = nn.Sequential(nn.Linear(I, 128), nn.Sigmoid(), nn.Linear(128, I), nn.Sigmoid())
    D = nn.Sequential(nn.Linear(I, 150), nn.Sigmoid(), nn.Linear(150, 150), nn.Sigmoid(), nn.Linear(150, 150), nn.Sigmoid(),
                      nn.Linear(150, 1), nn.Sigmoid())
    for k, v in att.netG_state().items():
        assert torch.equal(v.cpu(), G.state_dict()[k[len("main."):]])
    for k, v in att.netD_state().items():
        assert torch.equal(v.cpu(), D.state_dict()[k[len("main."):]])
    with pytest.raises(Exception):
        model.from_config("attacker", "aush", device=torch.device("cpu"), **kw).I(dataset=StubExplicit(mat, batch, dev))


def test_aush_over_the_b200_explicit_dataset_matches_the_reference():
    """Same golden run, batches from recad_b200.explicit.ExplicitData (device-side lazy `users_mat`): the epoch losses and
    the generator state must not depend on which dataset class hands out the rows."""
    from recad_b200 import dataset, model
    dev = torch.device("cuda:0")
    mat, G, D, kw, batch, targets, z = util.aush_case("b")
    tr, te = z["train"].astype(np.float64), z["test"].astype(np.float64)
    ds = dataset.from_config("explicit", "dev", device=dev, batch_size=batch, train_dict=tr.copy(), valid_dict=te.copy(), test_dict=te.copy())
    assert np.array_equal(ds.train_mat, mat)
    np.random.seed(1)
    b0 = next(iter(ds.generate_batch()))
    assert b0["users"].is_cuda and not dict.__contains__(b0, "users_mat")
    assert np.array_equal(b0["users_mat"].cpu().numpy(), mat[b0["users"].cpu().numpy()])
    att = model.from_config("attacker", "aush", device=dev, **kw).I(dataset=ds)
    att.load_netG_state(G)
    att.load_netD_state(D)
    np.random.set_state(("MT19937", z["b_np_key_start"], int(z["b_np_pos_start"]), 0, 0.0))
    for e, gold in enumerate(z["b_losses"]):
        assert np.allclose(att.train_step(target_id_list=targets), gold, rtol=1e-4, atol=0), e
    st = np.random.get_state()
    assert np.array_equal(st[1], z["b_np_key_mid"]) and st[2] == int(z["b_np_pos_mid"])


def test_aush_without_selected_items_matches_the_oracle():
    """selected_ids = []: no generated column, empty ZR pool (no shuffle draws), the discriminator still trains on the fillers."""
    from oracle import aush as oa
    from recad_b200 import model
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(9)
    mat = ((rng.random((150, 260)) < 0.15) * rng.integers(1, 6, (150, 260))).astype(np.float32)
    kw = dict(selected_ids=[], filler_num=9, attack_num=5, ZR_ratio=0.2)
    torch.manual_seed(1)
    att = model.from_config("attacker", "aush", device=dev, **kw).I(dataset=StubExplicit(mat, 64, dev))
    G0, D0 = {k: v.cpu().numpy() for k, v in att.netG_state().items()}, {k: v.cpu().numpy() for k, v in att.netD_state().items()}
    o = oa.AushOracle(mat, G0, D0, batch_size=64, **kw)
    np.random.seed(2)
    mine = [att.train_step(target_id_list=[4]) for _ in range(2)]
    s1 = np.random.get_state()
    np.random.seed(2)
    ref = [o.train_step([4]) for _ in range(2)]
    s2 = np.random.get_state()
    assert s1[2] == s2[2] and np.array_equal(s1[1], s2[1])
    for a, b in zip(mine, ref):
        assert np.allclose(a[0], b[0], rtol=1e-4) and np.allclose(a[3], b[3], rtol=1e-4) and a[1] == b[1] == 0.0 and a[2] == b[2] == 0.0
    np.random.seed(3)
    fake = att.generate_fake(target_id_list=[4])
    np.random.seed(3)
    assert np.array_equal(fake, o.generate_fake([4]))
