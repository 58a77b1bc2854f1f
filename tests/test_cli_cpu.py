"""CPU-side checks of the command-line driver: flag names / defaults of the reference, result-file layout."""
import numpy as np
import pytest
import torch


def test_flags_and_defaults_follow_the_reference():
    from sky_embeddings_b200 import search
    a = search.parse_arguments(["mim_1"])
    # similarity_search.py:22-57
    assert a.model_name == "mim_1" and a.target_indices == "[1,2]" and a.augment_targets == "True"
    assert a.max_pool == "True" and a.cls_token == "False" and a.snr_range == "[2,7]" and a.batch_size == 64
    assert a.metric == "cosine" and a.combine == "min" and a.display_channel == 2 and a.n_plot == 36 and a.n_save == 300
    assert a.target_fn.endswith(".h5") and a.test_fn.endswith(".h5") and a.data_dir is None and a.test_dirs is None
    b = search.parse_arguments(["m", "-tgt_i", "None", "-mp", "False", "-ct", "True", "-aug", "False", "-snr", "[1,9]",
                                "-bs", "512", "-m", "MSE", "-c", "mean", "-ns", "10", "-dd", "/x", "-tst_dirs", "a", "b"])
    assert (b.target_indices, b.max_pool, b.cls_token, b.augment_targets) == ("None", "False", "True", "False")
    assert (b.snr_range, b.batch_size, b.metric, b.combine, b.n_save, b.data_dir) == ("[1,9]", 512, "MSE", "mean", 10, "/x")
    assert b.test_dirs == ["a", "b"]
    assert search.str2bool("True") and search.str2bool("t") and not search.str2bool("False")


def test_save_results_writes_the_reference_keys(tmp_path):
    from sky_embeddings_b200.similarity import save_results
    k, T = 7, 3
    arrs = dict(test_ra_decs=torch.rand(k, 2), test_scores=torch.rand(k), target_images=torch.rand(T, 5, 8, 8),
                target_latent=torch.rand(T, 5, 16), test_images=torch.rand(k, 5, 8, 8), test_latent=torch.rand(k, 5, 16))
    p = str(tmp_path / "r.npz")
    save_results(p, arrs["test_ra_decs"], arrs["test_scores"], arrs["target_images"], arrs["target_latent"],
                 arrs["test_images"], arrs["test_latent"])
    r = np.load(p)
    # similarity_search.py:178-181
    assert sorted(r.files) == sorted(["test_ra_decs", "test_scores", "target_images", "target_features", "test_images", "test_features"])
    assert np.array_equal(r["target_features"], arrs["target_latent"].numpy()) and np.array_equal(r["test_features"], arrs["test_latent"].numpy())
    assert np.array_equal(r["test_scores"], arrs["test_scores"].numpy())


def test_cli_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sky_embeddings_b200 import search
    with pytest.raises(SystemExit):
        search.main(["m", "--encoder", "tests.test_gpu_cli:stub_factory"])


def test_resident_simsearch_returns_the_reference_tuple_with_a_model(monkeypatch):
    """resident_simsearch(model=...) re-encodes the winners as utils/similarity.py:124-130 does; checked on the CPU with a
    stand-in bank and a torch stand-in for the token-select kernel (the search itself is covered by the GPU tests):
    shapes, order, padding rows, wrapped (.module) models."""
    from sky_embeddings_b200 import feeder
    from sky_embeddings_b200.feeder import resident_simsearch
    monkeypatch.setattr(feeder, "select_tokens", lambda lat, n_extra, cls_token, max_pool:
                        lat[:, n_extra:].max(1, keepdim=True)[0] if max_pool else lat[:, n_extra:])
    n, D, P, k = 12, 8, 4, 5
    g = torch.Generator().manual_seed(3)
    samples = torch.rand(n, 5, 8, 8, generator=g)
    ra_decs = torch.rand(n, 2, generator=g)
    target = torch.rand(6, 1 + P, D, generator=g)
    want_idx = torch.tensor([7, 2, 9, -1, -1])           # a bank shorter than n_save pads with -1 (+-inf scores)
    want_sc = torch.tensor([0.9, 0.8, 0.1, float("-inf"), float("-inf")])

    class StubBank:
        device = torch.device("cpu")

        def query_from_targets(self, tsel, use_weights):
            assert tsel.shape == (6, 1, D)               # max_pool=True: one token per target row
            return tsel.mean((0, 1)), torch.full((D,), 1.0 / D)

        def search(self, t, w, k, metric, combine):
            assert k == 5 and metric == "cosine" and combine == "min" and w is not None
            return want_sc[None], want_idx[None]

    class StubEncoder:
        num_extra_tokens = 1

        def forward_features(self, x, ra_dec=None, reshape_out=False):
            assert ra_dec.shape == (x.shape[0], 2) and reshape_out is False
            lat = x.reshape(x.shape[0], -1)[:, : (1 + P) * D].reshape(x.shape[0], 1 + P, D)
            return lat, None, None

    class Wrapped:
        def __init__(self, m):
            self.module = m

    for model in (StubEncoder(), Wrapped(StubEncoder())):
        bs, bl, bra, sc = resident_simsearch(StubBank(), target, ra_decs, samples, num_extra_tokens=1, n_save=k,
                                             max_pool=True, model=model)
        assert bs.shape == (k, 5, 8, 8) and bl.shape == (k, 1 + P, D) and bra.shape == (k, 2) and torch.equal(sc, want_sc)
        assert torch.equal(bs[:3], samples[[7, 2, 9]]) and torch.equal(bra[:3], ra_decs[[7, 2, 9]])
        assert torch.count_nonzero(bs[3:]) == 0 and torch.count_nonzero(bra[3:]) == 0
        assert torch.equal(bl[:3], samples[[7, 2, 9]].reshape(3, -1)[:, : (1 + P) * D].reshape(3, 1 + P, D))
    # without a model: the winners' indices, as before
    bs, idx, bra, sc = resident_simsearch(StubBank(), target, ra_decs, samples, num_extra_tokens=1, n_save=k, max_pool=True)
    assert torch.equal(idx, want_idx)
    with pytest.raises(ValueError):
        resident_simsearch(StubBank(), target, ra_decs, None, num_extra_tokens=1, n_save=k, max_pool=True, model=StubEncoder())
