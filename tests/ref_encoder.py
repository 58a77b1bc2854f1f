"""The reference's OWN `mim_1` encoder (utils/mim_vit.py, unmodified) for BASELINE config 1, built with random-init
weights: utils/mim_vit.py is imported from /root/reference (build container) or from the byte-identical copy in
oracle/_ref (GPU box; recipe and sha256 manifest in oracle/ref_harness.py), behind the test-only timm shim in
tests/shims.  configs/mim_1.ini lacks the `attn_pool` / `ra_dec` keys build_model indexes (utils/mim_vit.py:31-32,
SURVEY.md appendix B): they are injected as False, which is what the released mim_1 checkpoint was trained with.
Test / golden infrastructure only."""
import configparser
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    from oracle.ref_harness import ref_path
    return ref_path("utils/mim_vit.py") is not None and ref_path("configs/mim_1.ini") is not None


def build_mim1(device="cpu", seed=0):
    """(model wrapped in nn.DataParallel as the reference does, config).  Weights: torch.manual_seed(seed) on the CPU
    generator, so the build container and the GPU box construct the same model."""
    from oracle.ref_harness import load_reference_module, ref_path
    shims = os.path.join(HERE, "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    utils_dir = os.path.dirname(ref_path("utils/mim_vit.py"))
    if utils_dir not in sys.path:
        sys.path.append(utils_dir)                      # mim_vit.py does `from pos_embed import ...` (:13-16)
    mv = load_reference_module("utils/mim_vit.py", "ref_mim_vit")
    config = configparser.ConfigParser()
    config.read(ref_path("configs/mim_1.ini"))
    config["ARCHITECTURE"]["attn_pool"] = "False"
    config["ARCHITECTURE"]["ra_dec"] = "False"
    torch.manual_seed(seed)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model, _losses, _it = mv.build_model(config, "/nonexistent/mim_1.pth.tar", "cpu", build_optimizer=False)
    model.eval()
    return model.to(device), config
