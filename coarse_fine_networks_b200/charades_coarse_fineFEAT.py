"""Coarse-stream Charades loader with the surface of the reference's charades_coarse_fineFEAT.py, and the reader / writer of
the on-disk fine-feature layout that extract_fineFEAT.py produces.

    save_fine_features(feat, save_dir, name)      extract_fineFEAT.py:172-173   one torch.save'd [1,C,Tf,7,7] fp32 tensor per
    load_fine_features(fine_feat, keys, vid)      charades_coarse_fineFEAT.py:83-88   video and layer: <dir>/<layer>/<video id>
    Charades(split_file, split, root, fine_feat, feature_keys, spatial_transform, task, frames, gamma_tau, crops)   130-206
    mt_collate_fn(batch)                          209-260   clips / labels padded to the longest item, features zero-padded and
                                                            CAPPED at 128 steps with a feat_mask

The joint two-stream step (train.coarse_fine_forward) hands the fine features to the coarse stream in memory; this module is
for the reference's two-phase workflow (features extracted once, coarse stream trained from disk) and for feature files
written by the reference.  Frame selection, labels, views and the draw order are inherited from charades_fine.Charades (the
reference duplicates that code in both loaders); pixels go through the GPU clip kernel.
"""
import os

import numpy as np
import torch

from . import charades_fine as _fine
from .charades_fine import make_dataset  # noqa: F401  (same function in both reference loaders, 91-127)

FEAT_CAP = 128          # charades_coarse_fineFEAT.py:211 ("NEW LIMIT FOR XYTC MIXING")


def save_fine_features(feat, save_dir, name):
    """feat: {layer: [1,C,Tf,7,7]} (x3d_fine global_tower output) -> <save_dir>/<layer>/<name>, CPU fp32 tensors."""
    for layer, t in feat.items():
        os.makedirs(os.path.join(save_dir, layer), exist_ok=True)
        torch.save(t.detach().to("cpu", torch.float32), os.path.join(save_dir, layer, name))


def extract_fine_features(fine_net, loader, save_dir, log=None):
    """The loop of extract_fineFEAT.py:152-173: every video of `loader` (batch size 1, items of charades_fine.mt_collate_fn:
    clips [1,N,3,T,S,S], labels, masks, names) through the fine stream built with global_tower=True, features written in the
    layout the coarse loader reads.  -> number of videos written."""
    fine_net.train(False)                                                 # extract_fineFEAT.py:137
    getattr(fine_net, "module", fine_net).aggregate_sub_bn_stats()        # :138-139 -- eval-mode BN reads bn.running_*
    done = 0
    with torch.no_grad():
        for clips, _labels, masks, names in loader:
            b, n = clips.shape[:2]
            feat, _ = fine_net([clips.reshape((b * n,) + tuple(clips.shape[2:])), masks])
            save_fine_features(feat, save_dir, names[0])
            done += 1
            if log is not None:
                log(done, names[0], {k: tuple(v.shape) for k, v in feat.items()})
    return done


def load_fine_features(fine_feat, feature_keys, vid):
    """-> {layer: numpy [C,Tf,h,w]}; the 'gx' entry (a CDF saved as a vector) is viewed as [1,Tf,1,1] like the reference does."""
    out = {}
    for layer in feature_keys:
        t = torch.load(os.path.join(fine_feat, layer, vid), weights_only=False).squeeze(0)
        out[layer] = (t.view(1, -1, 1, 1) if layer == "gx" else t).numpy()
    return out


class Charades(_fine.Charades):
    """charades_coarse_fineFEAT.py:130-206: the fine loader's item plus the video's fine features, meta and duration.
    (`split` is used as given: this loader has no extract_feat switch.)"""

    def __init__(self, split_file, split, root, fine_feat, feature_keys, spatial_transform=None, task="class", frames=80,
                 gamma_tau=5, crops=1, device="cuda", cache=True, decode="pil", host_items=False):
        super().__init__(split_file, split, root, spatial_transform, task=task, frames=frames, gamma_tau=gamma_tau, crops=crops,
                         extract_feat=False, device=device, cache=cache, decode=decode, host_items=host_items)
        self.fine_feat, self.feature_keys = fine_feat, list(feature_keys)

    def host_item(self, index):
        s = self.sample(index)                                            # window draw first ...
        s["feat"] = load_fine_features(self.fine_feat, self.feature_keys, s["vid"])
        self.spatial_transform.randomize_parameters(224)                  # ... then the transform's draws (reference order)
        s["tstate"] = self.spatial_transform.get_state()
        return s

    def finish(self, s):
        """-> (clips [N,3,T,S,S] on `device`, label, feat {layer: numpy [C,Tf,7,7]}, meta int64 [4], vid, duration)."""
        self.spatial_transform.set_state(s["tstate"])
        clips, label = self.views(self.spatial_transform.clip(self.frames_on_device(s)), s["label"], s["frame_count"])
        return clips, label, s["feat"], s["meta"], s["vid"], self.data[s["index"]][2]

    def device_collate(self, batch):
        return mt_collate_fn([self.finish(s) for s in batch])


def mt_collate_fn(batch):
    """charades_coarse_fineFEAT.py:209-260 -> [clips [B,N,3,Tmax,S,S], labels [B,C,TLmax], masks [B,TLmax],
    feat {layer: [B,C,Tf',h,w]}, feat_masks [B,Tf'], meta [B,4], vids, durations float64 [B]] with Tf' = min(longest, 128)."""
    B = len(batch)
    t_clip = max(item[0].shape[2] for item in batch)
    t_label = max(item[1].shape[1] for item in batch)
    first_key = next(iter(batch[0][2]))
    t_feat = min(max(item[2][first_key].shape[1] for item in batch), FEAT_CAP)
    ref_clip = batch[0][0]
    clips = ref_clip.new_zeros((B,) + tuple(ref_clip.shape[:2]) + (t_clip,) + tuple(ref_clip.shape[3:]), dtype=torch.float32)
    labels = torch.zeros(B, batch[0][1].shape[0], t_label)
    masks = torch.zeros(B, t_label)
    feat_masks = torch.zeros(B, t_feat)
    feat = {}
    for layer, f0 in batch[0][2].items():
        c, _, h, w = f0.shape
        feat[layer] = torch.zeros(B, c, t_feat, h, w)
    for i, (clip, label, item_feat, _meta, _vid, _dur) in enumerate(batch):
        clips[i, :, :, :clip.shape[2]] = clip
        labels[i, :, :label.shape[1]] = label
        masks[i, :label.shape[1]] = 1
        feat_masks[i, :min(FEAT_CAP, item_feat[first_key].shape[1])] = 1
        for layer, f in item_feat.items():
            keep = min(FEAT_CAP, f.shape[1])
            feat[layer][i, :, :keep] = torch.as_tensor(f[:, :keep])
    meta = torch.stack([torch.as_tensor(item[3]) for item in batch], 0)
    durations = torch.tensor([float(item[5]) for item in batch], dtype=torch.float64)
    return [clips, labels, masks, feat, feat_masks, meta, tuple(item[4] for item in batch), durations]
