"""Fine-stream Charades loader mirror (coarse_fine_networks_b200/charades_fine.py) on CPU against the outputs of the reference's
own loader (tests/golden/charades_loader.npz: charades_fine.py make_dataset / Charades.__getitem__ / mt_collate_fn run on two
synthetic JPEG videos by tests/golden/make_golden.py).  The JPEG files are rebuilt from the golden; pixels are processed by
a test double of spatial_transforms.Compose whose clip() calls the numpy oracle instead of the CUDA kernel (the kernel itself
is checked bit-exactly against the same oracle in tests/test_clip_gpu.py), so every other statement of the loader -- video
selection, labels, random start frame, frame indices, transform draws, multi-view slicing, padding -- runs as shipped."""
import hashlib
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import clip_oracle as CO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "charades_loader.npz")
SCALES = [224 / 256., 224 / 320.]


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    from coarse_fine_networks_b200 import charades_fine as L
    from coarse_fine_networks_b200 import spatial_transforms as ST

    class OracleCompose(ST.Compose):
        def clip(self, frames, out=None, t_pad=0):
            f = frames.cpu().numpy()
            x1, y1, crop, S, flip = self.params(f.shape[2], f.shape[1])
            return torch.from_numpy(CO.clip_preprocess(f, (x1, y1, crop), S, flip, self._norm.mean, self._norm.std, t_pad))

    g = np.load(GOLD)
    tmp = tmp_path_factory.mktemp("charades")
    root = os.path.join(tmp, "frames")
    blob, off = g["jpeg_bytes"], 0
    for name, size in zip(g["jpeg_names"], g["jpeg_sizes"]):
        path = os.path.join(root, str(name))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(blob[off:off + int(size)].tobytes())
        off += int(size)
    split_file = os.path.join(tmp, "split.json")
    with open(split_file, "w") as f:
        f.write(str(g["split_json"]))
    mean, std = list(g["mean"]), list(g["std"])
    train_tr = OracleCompose([ST.MultiScaleRandomCropMultigrid(SCALES, 224), ST.RandomHorizontalFlip(), ST.ToTensor(255),
                              ST.Normalize(mean, std)])
    val_tr = OracleCompose([ST.CenterCropScaled(224), ST.ToTensor(255), ST.Normalize(mean, std)])
    return type("E", (), dict(L=L, g=g, root=root, split_file=split_file, train_tr=train_tr, val_tr=val_tr))


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()


def check(env, name, clips, label, vid):
    g = env.g
    assert tuple(clips.shape) == tuple(g[name + "/shape"]), name
    assert np.array_equal(clips[..., ::37, ::41].numpy().view(np.uint32), g[name + "/sample"].view(np.uint32)), name
    assert sha(clips) == str(g[name + "/sha256"]), name
    assert np.array_equal(label.numpy(), g[name + "/label"]) and vid == str(g[name + "/vid"]), name


def test_make_dataset_selection_and_labels(env):
    ds = env.L.make_dataset(env.split_file, "training", env.root, cache=False)
    assert [d[0] for d in ds] == [str(v) for v in env.g["dataset_vids"]]           # SHORT (< 162 frames) and OTHER (subset) dropped
    assert np.array_equal(ds[1][1], env.g["dataset_label_VIDB"]) and ds[1][3] == 185
    # the cache file is written in a form the reference's np.load(allow_pickle=True) path reads back
    ds2 = env.L.make_dataset(env.split_file, "training", env.root, cache=True)
    ds3 = env.L.make_dataset(env.split_file, "training", env.root, cache=True)
    assert os.path.exists(env.split_file[:-5] + "_traininglabeldata_160.npy")
    assert all(a[0] == b[0] and np.array_equal(a[1], b[1]) and a[3] == b[3] for a, b in zip(ds2, ds3))
    os.remove(env.split_file[:-5] + "_traininglabeldata_160.npy")


def test_training_samples_same_seed_same_clip(env):
    ds = env.L.Charades(env.split_file, "training", env.root, env.train_tr, task="class", frames=80, gamma_tau=5, crops=1,
                        device="cpu", cache=False)
    for seed in (3, 12):
        random.seed(seed)
        check(env, f"train_class_seed{seed}", *ds[seed % 2])


@pytest.mark.parametrize("name,task,crops,idx", [("test_loc_c1", "loc", 1, 1), ("test_loc_c2", "loc", 2, 1),
                                                 ("test_class_c2", "class", 2, 1), ("test_loc_c1_a", "loc", 1, 0)])
def test_testing_views(env, name, task, crops, idx):
    dv = env.L.Charades(env.split_file, "training", env.root, env.val_tr, task=task, frames=80, gamma_tau=5, crops=crops,
                        extract_feat=True, device="cpu", cache=False)
    random.seed(0)
    check(env, name, *dv[idx])


def test_sample_meta_and_indices(env):
    ds = env.L.Charades(env.split_file, "training", env.root, env.train_tr, task="loc", device="cpu", cache=False)
    random.seed(5)
    s = ds.sample(1)
    random.seed(5)
    start = random.randint(1, max(10, 185 - 160))
    assert s["start_f"] == start and s["frame_count"] == 160 and s["stride_f"] == 10 and s["frames"].shape == (16, 48, 64, 3)
    assert s["meta"].tolist() == [start // 10, 16, 18, 1] and tuple(s["label"].shape) == (157, 160)


def test_collate_padding(env):
    dv = env.L.Charades(env.split_file, "training", env.root, env.val_tr, task="loc", frames=80, gamma_tau=5, crops=1,
                        extract_feat=True, device="cpu", cache=False)
    b = env.L.mt_collate_fn([dv[0], dv[1]])
    g = env.g
    assert tuple(b[0].shape) == tuple(g["collate/shape"]) and sha(b[0]) == str(g["collate/sha256"])
    assert np.array_equal(b[1].numpy(), g["collate/labels"]) and np.array_equal(b[2].numpy(), g["collate/masks"])
    assert list(b[3]) == [str(v) for v in g["collate/vids"]]


# ---------------------------------------------------------------------------------------------------------------------------
# coarse-stream loader + the on-disk fine-feature layout (charades_coarse_fineFEAT.py, extract_fineFEAT.py:172-173)
# ---------------------------------------------------------------------------------------------------------------------------
FKEYS = ["layer1", "conv5"]


@pytest.fixture(scope="module")
def coarse_items(env, tmp_path_factory):
    from coarse_fine_networks_b200 import charades_coarse_fineFEAT as LC
    g = env.g
    feat_dir = str(tmp_path_factory.mktemp("feat"))
    for vid in ("VIDA", "VIDB"):                               # written with OUR writer, in the reference's layout
        LC.save_fine_features({k: torch.from_numpy(g[f"feat_file/{k}/{vid}"]) for k in FKEYS}, feat_dir, vid)
    dc = LC.Charades(env.split_file, "training", env.root, feat_dir, FKEYS, env.train_tr, task="loc", frames=80, gamma_tau=5,
                     crops=1, device="cpu", cache=False)
    random.seed(21)
    return LC, feat_dir, [dc[0], dc[1]]


def test_feature_files_round_trip_in_reference_layout(env, coarse_items):
    LC, feat_dir, _ = coarse_items
    for vid in ("VIDA", "VIDB"):
        assert sorted(os.listdir(feat_dir)) == sorted(FKEYS) and os.path.isfile(os.path.join(feat_dir, "conv5", vid))
        raw = torch.load(os.path.join(feat_dir, "layer1", vid), weights_only=False)      # what the reference's loader does
        assert raw.dtype == torch.float32 and raw.dim() == 5 and raw.shape[0] == 1
        back = LC.load_fine_features(feat_dir, FKEYS, vid)
        for k in FKEYS:
            assert np.array_equal(back[k], env.g[f"feat_file/{k}/{vid}"][0])


def test_coarse_items_match_reference(env, coarse_items):
    _, _, items = coarse_items
    g = env.g
    for i, (clips, label, feat, meta, vid, dur) in enumerate(items):
        check(env, f"coarse_item{i}", clips, label, vid)
        assert meta.tolist() == g[f"coarse_item{i}/meta"].tolist() and float(dur) == float(g[f"coarse_item{i}/dur"])
        for k in FKEYS:
            assert tuple(feat[k].shape) == tuple(g[f"coarse_item{i}/feat_shape/{k}"])
            assert sha(torch.from_numpy(feat[k])) == str(g[f"coarse_item{i}/feat_sha256/{k}"])


def test_coarse_collate_caps_features_at_128(env, coarse_items):
    LC, _, items = coarse_items
    g = env.g
    b = LC.mt_collate_fn(items)
    assert tuple(b[0].shape) == tuple(g["ccollate/clips_shape"]) and sha(b[0]) == str(g["ccollate/clips_sha256"])
    assert np.array_equal(b[1].numpy(), g["ccollate/labels"]) and np.array_equal(b[2].numpy(), g["ccollate/masks"])
    assert np.array_equal(b[4].numpy(), g["ccollate/feat_masks"]) and b[4].shape[1] == 128
    assert np.array_equal(b[5].numpy(), g["ccollate/meta"]) and list(b[6]) == [str(v) for v in g["ccollate/vids"]]
    assert np.array_equal(b[7].numpy(), g["ccollate/dur"]) and b[7].dtype == torch.float64
    for k in FKEYS:
        assert tuple(b[3][k].shape) == tuple(g[f"ccollate/feat_shape/{k}"]) and sha(b[3][k]) == str(g[f"ccollate/feat_sha256/{k}"])


def test_extract_loop_writes_what_the_coarse_loader_reads(env, tmp_path):
    """extract_fine_features (extract_fineFEAT.py:152-173) with a stand-in for the fine stream: one file per layer and video,
    batch and view dimensions merged before the call, read back by load_fine_features."""
    from coarse_fine_networks_b200 import charades_coarse_fineFEAT as LC
    seen = []

    class Tower(torch.nn.Module):
        aggregated = 0

        def aggregate_sub_bn_stats(self):              # extract_fineFEAT.py:138-139: called once, right after train(False)
            assert not self.training
            self.aggregated += 1

        def forward(self, inp):
            x, masks = inp
            seen.append((tuple(x.shape), self.training, torch.is_grad_enabled()))
            t = x.shape[2]
            return {"layer1": x.mean(dim=(3, 4), keepdim=True)[:, :2].expand(-1, -1, t, 7, 7) + 0.0,
                    "conv5": torch.ones(x.shape[0], 3, t, 7, 7)}, masks

    dv = env.L.Charades(env.split_file, "training", env.root, env.val_tr, task="loc", frames=80, gamma_tau=5, crops=1,
                        extract_feat=True, device="cpu", cache=False)
    loader = [env.L.mt_collate_fn([dv[i]]) for i in range(2)]
    net = Tower().train()
    n = LC.extract_fine_features(net, loader, str(tmp_path))
    assert net.aggregated == 1
    assert n == 2 and [s[0][:3] for s in seen] == [(1, 3, 17), (1, 3, 18)] and all(not s[1] and not s[2] for s in seen)
    for vid, t in (("VIDA", 17), ("VIDB", 18)):
        f = LC.load_fine_features(str(tmp_path), ["layer1", "conv5"], vid)
        assert f["layer1"].shape == (2, t, 7, 7) and f["conv5"].shape == (3, t, 7, 7) and f["conv5"].dtype == np.float32
