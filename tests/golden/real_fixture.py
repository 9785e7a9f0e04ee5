"""The reference's REAL FB15k-237 fixture (data/fb15k237/mapped_to_ids: valid.txt, test.txt and the six id / token maps;
its train.txt is absent from the checkout) staged as a three-split dataset directory.

Shared by ``make_golden.py`` (which feeds the staged directory to the unmodified reference) and by the tests (which feed
the same directory to this repo's dataset build), so both sides read byte-identical files:

    train.txt := the fixture's valid.txt                 (17,535 triples)
    valid.txt := the first 10,000 lines of its test.txt
    test.txt  := the remaining lines of its test.txt     (10,466 triples)

``fb15k237_ids.tar.gz`` next to this file holds the eight fixture files verbatim (data, not code)."""
import os
import shutil
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "fb15k237_ids.tar.gz")
MAPS = ("entity_id_map.txt", "relation_id_map.txt", "entity_token_id_map.txt", "relation_token_id_map.txt",
        "entity_id_tokens_ids_map.txt", "relation_id_tokens_ids_map.txt")
FILES = MAPS + ("valid.txt", "test.txt")
N_VALID = 10_000


def make_archive(src_dir: str) -> None:
    with tarfile.open(ARCHIVE, "w:gz") as tar:
        for f in FILES:
            tar.add(os.path.join(src_dir, f), arcname=f)


def stage(root: str, src_dir: str = None) -> str:
    """Writes the three-split directory into ``root`` from ``src_dir`` (the reference checkout) or from the archive."""
    os.makedirs(root, exist_ok=True)
    raw = os.path.join(root, "_raw")
    os.makedirs(raw, exist_ok=True)
    if src_dir is not None:
        for f in FILES:
            shutil.copy(os.path.join(src_dir, f), raw)
    else:
        with tarfile.open(ARCHIVE, "r:gz") as tar:
            for m in tar.getmembers():
                if m.name not in FILES or not m.isfile():
                    raise ValueError(f"unexpected member {m.name!r} in {ARCHIVE}")
                with tar.extractfile(m) as src, open(os.path.join(raw, m.name), "wb") as dst:
                    shutil.copyfileobj(src, dst)
    for f in MAPS:
        shutil.copy(os.path.join(raw, f), os.path.join(root, f))
    shutil.copy(os.path.join(raw, "valid.txt"), os.path.join(root, "train.txt"))
    with open(os.path.join(raw, "test.txt"), "rb") as f:
        lines = f.readlines()
    with open(os.path.join(root, "valid.txt"), "wb") as f:
        f.writelines(lines[:N_VALID])
    with open(os.path.join(root, "test.txt"), "wb") as f:
        f.writelines(lines[N_VALID:])
    shutil.rmtree(raw)
    return root


def seeded_weights(state_dict_items, seed: int = 7, std: float = 0.1):
    """Deterministic weights for both sides: float tensors of the state dict, in sorted key order, drawn from
    numpy's default_rng(seed) (platform-stable) as standard_normal * std. Returns {key: float32 array}."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = {}
    for k, shape in sorted(state_dict_items):
        out[k] = (rng.standard_normal(shape) * std).astype(np.float32)
    return out
