"""Command-line driver: the reference's ``similarity_search.py`` / ``sky_sim_search.py`` on the B200 engine.

    python -m sky_embeddings_b200.search MODEL_NAME [-tgt_fn F.h5] [-tst_fn F.h5 | -tst_dirs DIR ...]
           [-tgt_i '[1,2]'] [-aug True] [-mp True] [-ct False] [-snr '[2,7]'] [-bs 64] [-m cosine] [-c min]
           [-dc 2] [-np 36] [-ns 300] [-dd DATA_DIR]

Flag names, defaults and meaning are the reference's (similarity_search.py:16-77, sky_sim_search.py:17-56); the
result file keeps its name and keys (similarity_search.py:178-181: ``results/{model}_{tgt_fn[:-3]}_simsearch_results_f.npz``
for an h5 bank; sky_sim_search.py:171-174: ``..._simsearch_results.npz`` for tiles; keys ``test_ra_decs``,
``test_scores``, ``target_images``, ``target_features``, ``test_images``, ``test_features``).

What runs where:
  * bank file -> S/N window -> ascending ``test_indices``: ``ingest.H5Cutouts.select_snr`` (CUDA S/N kernel)
    [similarity_search.py:122-130];
  * targets -> ``feeder.mae_latent`` with the reference's target augmentation (64 copies) [:159-162];
  * search -> the drop-in ``similarity.mae_simsearch`` (exact reference semantics incl. first-batch statistics and the
    winners' re-encode) [:169-171]; ``--resident`` encodes the bank once into a device-resident bank and searches it
    in one pass (``feeder.bank_from_loader`` + ``resident_simsearch``), the route to use when several target groups
    share one bank;
  * figures need matplotlib (absent here): skipped with a note.

The encoder is the reference's own PyTorch model and is NOT part of this package: by default
``utils.mim_vit.build_model`` / ``utils.vit.build_model`` are imported from ``--reference-root`` exactly as the
reference drivers do; ``--encoder pkg.module:factory`` plugs in any object with the duck-typed contract of
utils/similarity.py:47-52 (factory(config, device) -> model).
"""
from __future__ import annotations

import argparse
import ast
import configparser
import importlib
import os
import sys

import numpy as np
import torch


def str2bool(v):                       # utils/misc.py:6-7
    return v.lower() in ("yes", "true", "t", "1")


def parse_arguments(argv=None):
    parser = argparse.ArgumentParser('Similarity searching.', add_help=True)
    parser.add_argument("model_name", help="Name of model.", type=str)
    parser.add_argument("-tgt_fn", "--target_fn", type=str, default='HSC_dud_dwarf_galaxy_calexp_GIRYZ7610_64.h5')
    parser.add_argument("-tst_fn", "--test_fn", type=str, default='HSC_dud_unknown_calexp_GIRYZ7610_64.h5')
    parser.add_argument("-tst_dirs", "--test_dirs", type=str, nargs='+', default=None,
                        help="sky_sim_search.py mode: directories of tiles instead of an h5 bank")
    parser.add_argument("-tgt_i", "--target_indices", default='[1,2]')
    parser.add_argument("-aug", "--augment_targets", type=str, default='True')
    parser.add_argument("-mp", "--max_pool", type=str, default='True')
    parser.add_argument("-ct", "--cls_token", type=str, default='False')
    parser.add_argument("-snr", "--snr_range", default='[2,7]')
    parser.add_argument("-bs", "--batch_size", type=int, default=64)
    parser.add_argument("-m", "--metric", type=str, default='cosine')
    parser.add_argument("-c", "--combine", type=str, default='min')
    parser.add_argument("-dc", "--display_channel", type=int, default=2)
    parser.add_argument("-np", "--n_plot", type=int, default=36)
    parser.add_argument("-ns", "--n_save", type=int, default=300)
    parser.add_argument("-dd", "--data_dir", help="Data directory if different from <root>/data/", type=str, default=None)
    # engine-side options (not in the reference)
    parser.add_argument("--reference-root", default=None, help="checkout of teaghan/sky_embeddings (configs/, models/, utils/)")
    parser.add_argument("--encoder", default=None, help="pkg.module:factory returning the model (default: the reference's build_model)")
    parser.add_argument("--results-dir", default=None)
    parser.add_argument("--resident", action="store_true", help="encode the bank once and search the resident bank")
    parser.add_argument("--bank-dtype", default="fp32", choices=["fp32", "bf16"])
    parser.add_argument("--overlap", type=float, default=0.4, help="tile mode: cutout overlap (sky_sim_search.py:150)")
    parser.add_argument("--device", default="cuda")
    return parser.parse_args(argv)


def build_encoder(args, config, root, device):
    """(model, mae_config).  Default: what the reference drivers do (similarity_search.py:96-120)."""
    if args.encoder:
        mod, _, fn = args.encoder.partition(":")
        model = getattr(importlib.import_module(mod), fn)(config, device)
        return model, config
    if root is None:
        raise SystemExit("no encoder: pass --reference-root (the reference's own model code) or --encoder module:factory")
    if root not in sys.path:
        sys.path.insert(0, root)
    model_dir = os.path.join(root, 'models/')
    config_dir = os.path.join(root, 'configs/')
    model_filename = os.path.join(model_dir, args.model_name + '.pth.tar')
    if 'pretained_mae' in config['TRAINING']:
        from utils.vit import build_model as build_vit  # noqa: PLC0415
        mae_name = config['TRAINING']['pretained_mae']
        if mae_name == 'None':
            mae_filename, mae_config = 'None', config
        else:
            mae_config = configparser.ConfigParser()
            mae_config.read(config_dir + mae_name + '.ini')
            mae_filename = os.path.join(model_dir, mae_name + '.pth.tar')
        model, _, _ = build_vit(config, mae_config, model_filename, mae_filename, device, build_optimizer=False)
        return model, mae_config
    from utils.mim_vit import build_model as build_mim  # noqa: PLC0415
    model, _, _ = build_mim(config, model_filename, device, build_optimizer=False)
    return model, config


def load_tiles(test_dirs, config=None):
    """Tile mode input.  Directories of `[calexp-]HSC-<band>-<tract>-<patch>.fits` files are grouped per patch and read
    like the reference does (find_HSC_bands / load_fits_bands, utils/dataloaders.py:330-437: bands and min_bands from the
    model's config, NaN planes for missing bands, ra / dec from the first band's TAN WCS) by the built-in FITS reader
    (`fitslite`: uncompressed image HDUs); `*.npy` tiles [C, H, W] are taken as they are (no WCS).  Tiles are opened
    lazily, one per loader item."""
    from . import fitslite  # noqa: PLC0415
    from .ingest import load_tile_npy  # noqa: PLC0415
    bands, min_bands, use_calexp = ['G', 'R', 'I', 'Z', 'Y'], 5, True
    if config is not None and config.has_section('DATA'):
        if 'bands' in config['DATA']:
            bands = ast.literal_eval(config['DATA']['bands'])
        if 'min_bands' in config['DATA']:
            min_bands = int(config['DATA']['min_bands'])
        if 'use_calexp' in config['DATA']:
            use_calexp = str2bool(config['DATA']['use_calexp'])
    groups = fitslite.find_tile_bands(test_dirs, bands, min_bands, use_calexp=use_calexp, verbose=1)
    npys = []
    for d in test_dirs:
        npys += sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".npy"))
    if not groups and not npys:
        raise SystemExit(f"no FITS patches with at least {min_bands} of {bands} and no .npy tiles under {test_dirs}")

    class LazyTiles:
        def __len__(self):
            return len(groups) + len(npys)

        def __getitem__(self, i):
            if i < len(groups):
                return fitslite.load_tile_bands(groups[i], return_wc=True)
            return load_tile_npy(npys[i - len(groups)])
    return LazyTiles()


def main(argv=None):
    args = parse_arguments(argv)
    from . import ingest  # noqa: PLC0415
    from .feeder import bank_from_loader, mae_latent, resident_simsearch  # noqa: PLC0415
    from .similarity import mae_simsearch, save_results  # noqa: PLC0415

    model_name, target_fn = args.model_name, args.target_fn
    target_indices = ast.literal_eval(args.target_indices) if args.target_indices != 'None' else None
    augment_targets, max_pool, cls_token = str2bool(args.augment_targets), str2bool(args.max_pool), str2bool(args.cls_token)
    snr_range = ast.literal_eval(args.snr_range)
    root = args.reference_root
    data_dir = args.data_dir or os.path.join(root or os.getcwd(), 'data/')
    results_dir = args.results_dir or os.path.join(root or os.getcwd(), 'results/')
    os.makedirs(results_dir, exist_ok=True)

    device = torch.device(args.device)
    if device.type != "cuda" or not torch.cuda.is_available():
        raise SystemExit("this driver needs a CUDA device: the search path has no CPU fallback")
    print(f'Using Torch version: {torch.__version__}')
    print(f'Using a {device} device with {torch.cuda.device_count()} GPU(s)')

    config = configparser.ConfigParser()
    if root is not None:
        config.read(os.path.join(root, 'configs/', model_name + '.ini'))
    model, _mae_config = build_encoder(args, config, root, device)
    img_size = int(config['ARCHITECTURE']['img_size']) if config.has_section('ARCHITECTURE') else 64

    tiles_mode = args.test_dirs is not None
    target_src = ingest.H5Cutouts(os.path.join(data_dir, target_fn), img_size=img_size)
    target_loader = target_src.loader(indices=target_indices, batch_size=args.batch_size)
    if tiles_mode:
        test_loader = ingest.TileLoader(load_tiles(args.test_dirs, config), batch_size=args.batch_size, img_size=img_size,
                                        overlap=args.overlap, device=device)
        test_src = None
    else:
        print('Estimating S/N for test dataset images...')
        test_src = ingest.H5Cutouts(os.path.join(data_dir, args.test_fn), img_size=img_size)
        test_indices, _ = test_src.select_snr(device, snr_range, n_central_pix=8, batch_size=5000)
        print(f'{len(test_indices)} of {len(test_src)} test images inside S/N range {snr_range}')
        test_loader = test_src.loader(indices=test_indices, batch_size=args.batch_size, device=device)

    target_latent, target_images = mae_latent(model, target_loader, device, return_images=True,
                                              apply_augmentations=augment_targets, num_augmentations=64, remove_cls=False)

    if args.resident and not tiles_mode:
        bank, ra_decs, samples = bank_from_loader(model, test_loader, device, max_pool=max_pool, cls_token=cls_token,
                                                  nested_batches=False, bank_dtype=args.bank_dtype, keep_samples=True)
        enc = model.module if hasattr(model, "module") else model
        test_images, _idx, test_ra_decs, test_scores = resident_simsearch(
            bank, target_latent, ra_decs, samples, num_extra_tokens=enc.num_extra_tokens, n_save=args.n_save,
            metric=args.metric, combine=args.combine, use_weights=True, max_pool=max_pool, cls_token=cls_token)
        test_images = test_images.to(device)
        with torch.no_grad():       # the winners' re-encode of utils/similarity.py:124-130
            test_latent, _, _ = enc.forward_features(test_images, ra_dec=test_ra_decs, reshape_out=False)
        bank.close()
    else:
        test_images, test_latent, test_ra_decs, test_scores = mae_simsearch(
            model, target_latent, test_loader, device, metric=args.metric, combine=args.combine, use_weights=True,
            max_pool=max_pool, cls_token=cls_token, nested_batches=tiles_mode, n_save=args.n_save,
            bank_dtype=args.bank_dtype)

    try:
        import matplotlib  # noqa: F401, PLC0415
        print('matplotlib found: figures are the reference\'s plotting_fns.display_images job (not part of the search path)')
    except ImportError:
        print('matplotlib is not installed: skipping the target / result figures')

    out = os.path.join(results_dir, f'{model_name}_{target_fn[:-3]}_simsearch_results{"" if tiles_mode else "_f"}.npz')
    save_results(out, test_ra_decs, test_scores, target_images, target_latent, test_images, test_latent)
    print(f'Saved {out}')
    target_src.close()
    if test_src is not None:
        test_src.close()
    return out


if __name__ == "__main__":
    main()
