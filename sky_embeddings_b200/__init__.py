"""sky_embeddings_b200 -- B200-native exact top-k similarity search for the search path of
teaghan/sky_embeddings (utils/similarity.py as driven by similarity_search.py / sky_sim_search.py).

    from sky_embeddings_b200.similarity import mae_simsearch, compute_similarity   # reference names
    from sky_embeddings_b200 import Bank                                           # resident bank

Everything numeric runs in libskysearch.so (hand-written sm_100a CUDA behind a C ABI,
include/sky_search.h); importing the search API without the built library raises.
"""
__all__ = ["Bank", "PixelBank", "bank_from_loader", "resident_simsearch", "merge_candidates", "ShardedBank", "sharded_search", "shard_range",
           "H5Cutouts", "H5CutoutLoader", "TileLoader"]


def __getattr__(name):
    if name in ("Bank", "PixelBank", "merge_candidates"):
        from . import engine
        return getattr(engine, name)
    if name in ("ShardedBank", "sharded_search", "shard_range"):
        from . import distributed
        return getattr(distributed, name)
    if name in ("bank_from_loader", "resident_simsearch"):
        from . import feeder
        return getattr(feeder, name)
    if name in ("H5Cutouts", "H5CutoutLoader", "TileLoader"):
        from . import ingest
        return getattr(ingest, name)
    raise AttributeError(name)
