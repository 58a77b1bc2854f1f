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
