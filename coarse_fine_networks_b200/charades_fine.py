"""Fine-stream Charades loader with the surface of the reference's charades_fine.py, feeding the GPU
clip kernel instead of per-image PIL transforms.

    make_dataset(split_file, split, root, num_classes=157)        charades_fine.py:83-122
    Charades(split_file, split, root, spatial_transform, task, frames, gamma_tau, crops, extract_feat)   125-202
    mt_collate_fn(batch)                                          205-229

Host logic (which videos, which frame indices, the per-frame label matrix, the multi-view slicing, the zero padding) follows
the reference statement by statement, including the order of the draws from Python's `random` (start frame first, then the
transform's parameters), so one seed selects the same clips on both sides.  What changes is where the pixels are processed:
the decoded frames of a video go to the GPU as ONE uint8 tensor [T,H,W,3] and `spatial_transform.clip()` (one launch of
cf_clip_preprocess) produces the normalised [3,T,S,S] clip -- bit-identical to `[transform(img) for img in imgs]` + stack +
permute (charades_fine.py:170-172).  JPEG decoding: `decode="pil"` keeps the reference's host decode (pil_loader, 22-26);
`decode="nvjpeg"` reads the files as bytes and decodes the whole clip on the GPU (jpeg.JpegDecoder, nvJPEG batched decode)
straight into the uint8 tensor the clip kernel reads.

DataLoader workers: CUDA cannot be used in forked workers and pinned-memory collation does not apply to CUDA tensors, so with
`num_workers > 0` build the dataset with `host_items=True`: `__getitem__` then does only the host half (window draw, file
reads / PIL decode, the transform's random draw) and returns a plain dict; pass `collate_fn=host_collate` (workers also
run the collate function, so it only gathers the dicts) and finish every batch in the MAIN process with
`dataset.device_collate(samples)` (GPU decode + clip kernel + the reference's zero-padding collate).  Without it
(`host_items=False`, the default) items are CUDA tensors as before and the loader must run with num_workers=0, pin_memory=False.

Differences, on purpose: the label cache `<split>_<split>labeldata_160.npy` is written as an object array (the reference's
`np.save(list_of_tuples)` raises on numpy >= 1.24) and read back the way the reference reads it; the accimage backend is not
used.  `Charades.sample()` is the host half of `__getitem__` (no GPU needed).
"""
import json
import os
import random

import numpy as np
import torch
from PIL import Image


def pil_loader(path):
    """charades_fine.py:22-26 -> uint8 [H,W,3]."""
    with open(path, "rb") as f:
        with Image.open(f) as img:
            return np.asarray(img.convert("RGB"))


def video_loader(video_dir_path, vid, frame_indices, image_loader=pil_loader):
    """charades_fine.py:46-56: frames <root>/<vid>/<vid>-<000001>.jpg, stopping at the first missing index."""
    video = []
    for i in frame_indices:
        image_path = os.path.join(video_dir_path, vid, vid + "-" + str(i).zfill(6) + ".jpg")
        if not os.path.exists(image_path):
            return video
        video.append(image_loader(image_path))
    return video


def load_rgb_frames(image_dir, vid, start, num, stride, loader=video_loader):
    """charades_fine.py:73-80."""
    return loader(image_dir, vid, list(range(start, start + num, stride)))


def bytes_loader(path):
    """The undecoded JPEG stream of a frame (for decode="nvjpeg")."""
    with open(path, "rb") as f:
        return f.read()


def make_dataset(split_file, split, root, num_classes=157, cache=True):
    """charades_fine.py:83-122 -> list of (vid, label float32 [num_classes, num_frames], duration, num_frames).
    label[c, fr] = 1 iff ann.start < fr / fps < ann.end, fps = num_frames / duration (same float64 operations)."""
    pre_data_file = split_file[:-5] + "_" + split + "labeldata_160.npy"
    if cache and os.path.exists(pre_data_file) and os.path.getsize(pre_data_file) > 0:
        return [tuple(d) for d in np.load(pre_data_file, allow_pickle=True)]
    with open(split_file, "r") as f:
        data = json.load(f)
    dataset = []
    for vid in data.keys():
        if data[vid]["subset"] != split:
            continue
        if not os.path.exists(os.path.join(root, vid)):
            continue
        num_frames = len(os.listdir(os.path.join(root, vid)))
        if num_frames < (2 * 80 + 2):
            continue
        label = np.zeros((num_classes, num_frames), np.float32)
        fps = num_frames / data[vid]["duration"]
        t = np.arange(num_frames) / fps
        for ann in data[vid]["actions"]:
            label[ann[0], (t > ann[1]) & (t < ann[2])] = 1
        dataset.append((vid, label, data[vid]["duration"], num_frames))
    if cache:
        arr = np.empty(len(dataset), dtype=object)
        for i, d in enumerate(dataset):
            arr[i] = d
        np.save(pre_data_file, arr, allow_pickle=True)
    return dataset


class Charades(torch.utils.data.Dataset):
    """charades_fine.py:125-202.  `spatial_transform` is a coarse_fine_networks_b200.spatial_transforms.Compose (or any
    object with randomize_parameters(c_size) and clip(frames_u8)); `device` is where the frames are processed."""

    def __init__(self, split_file, split, root, spatial_transform=None, task="class", frames=80, gamma_tau=5, crops=1,
                 extract_feat=False, device="cuda", cache=True, decode="pil", host_items=False):
        if decode not in ("pil", "nvjpeg"):
            raise ValueError("decode must be 'pil' (host, as the reference) or 'nvjpeg' (GPU)")
        self.decode, self.host_items, self._decoder = decode, host_items, None
        self.data = make_dataset(split_file, split, root, cache=cache)
        self.split_file = split_file
        self.root = root
        self.frames = frames * 2
        self.gamma_tau = gamma_tau * 2
        self.spatial_transform = spatial_transform
        self.crops = crops
        self.split = "testing" if extract_feat else split
        self.task = task
        self.device = device

    def __len__(self):
        return len(self.data)

    def plan(self, nf):
        """Which frames of an nf-frame video one item reads (charades_fine.py:147-166): -> (start_f, frame_count, stride_f).
        Testing mode reads the whole video from frame 1 (at 1/crops of the stride for localisation); training mode draws the
        start of a self.frames-long window with random.randint -- the first draw of the item, before the transform's."""
        testing = self.split == "testing"
        count = nf if testing else min(self.frames, nf)
        start = 1 if testing else random.randint(1, max(self.gamma_tau, nf - count))
        stride = self.gamma_tau // self.crops if (testing and self.task == "loc") else self.gamma_tau
        return start, count, stride

    def sample(self, index):
        """Host half of __getitem__ (no GPU): draws the window, decodes its frames, cuts the labels to it.
        -> dict(frames uint8 [T,H,W,3], label, vid, frame_count, start_f, stride_f, meta int64 [4])."""
        vid, full_label, _, nf = self.data[index]
        start, count, stride = self.plan(nf)
        if self.decode == "nvjpeg":                                       # undecoded streams: the GPU decodes the whole clip at once
            decoded = video_loader(self.root, vid, list(range(start, start + count, stride)), image_loader=bytes_loader)
        else:
            decoded = load_rgb_frames(self.root, vid, start, count, stride)
        window = torch.from_numpy(np.ascontiguousarray(full_label[:, start - 1:start - 1 + count]))
        if self.task == "class":
            window = window.max(dim=1).values                             # clip-level label: any frame positive
        g = self.gamma_tau
        meta = torch.tensor([start // g, count // g, nf // g, stride // g], dtype=torch.int64)      # charades_fine.py:193-194
        frames = list(decoded) if self.decode == "nvjpeg" else np.stack(decoded, 0)
        return dict(frames=frames, label=window, vid=vid, frame_count=count, start_f=start, stride_f=stride, meta=meta, index=index)

    def view_indices(self, n_decoded, frame_count):
        """Frame positions (into the decoded clip) of every view of the item (charades_fine.py:174-191): one view holding all
        frames in training; in testing `crops` views -- evenly spaced windows of self.frames // gamma_tau frames for
        classification, interleaved sub-samplings (i, i + crops, ...) cut to frame_count // gamma_tau for localisation."""
        everything = [list(range(n_decoded))]
        if self.split != "testing":
            return everything
        if self.task == "class":
            per_view = self.frames // self.gamma_tau
            step = int((n_decoded - 1 - per_view) // (self.crops - 1))
            starts = [0] * self.crops if step == 0 else list(range(0, step * self.crops, step))
            return [list(range(s0, min(s0 + per_view, n_decoded))) for s0 in starts]
        if self.task == "loc":
            keep = frame_count // self.gamma_tau
            return [list(range(i, n_decoded, self.crops))[:keep] for i in range(self.crops)]
        return everything

    def views(self, imgs_l, label, frame_count):
        """[3,T,S,S] -> clips [N,3,T',S,S]; localisation labels are cut to whole strides (charades_fine.py:189)."""
        picks = self.view_indices(imgs_l.shape[1], frame_count)
        clips = torch.stack([imgs_l[:, torch.as_tensor(p, device=imgs_l.device)] for p in picks], 0)
        if self.split == "testing" and self.task == "loc":
            label = label[:, :(frame_count // self.gamma_tau) * self.gamma_tau]
        return clips, label

    def frames_on_device(self, s):
        """Decoded frames of a sample as ONE uint8 CUDA tensor [T,H,W,3]: host->device copy of PIL's output, or the nvJPEG
        batched decode of the undecoded streams."""
        if isinstance(s["frames"], list):
            if self._decoder is None:
                from .jpeg import JpegDecoder
                self._decoder = JpegDecoder()
            return self._decoder.decode(s["frames"], device=self.device)
        return torch.from_numpy(s["frames"]).to(self.device, non_blocking=True)

    def host_item(self, index):
        """Everything of __getitem__ that needs no GPU, with the reference's `random` draw order: the window first
        (sample), then the transform's parameters (charades_fine.py:147-170) -- shipped as plain numbers."""
        s = self.sample(index)
        self.spatial_transform.randomize_parameters(224)                  # charades_fine.py:170 (224 is hard-coded there)
        s["tstate"] = self.spatial_transform.get_state()
        return s

    def finish(self, s):
        """Device half of an item: decode / upload -> clip kernel (171-172 in one launch) -> views."""
        self.spatial_transform.set_state(s["tstate"])
        imgs_l = self.spatial_transform.clip(self.frames_on_device(s))    # [3,T,S,S]
        clips, label = self.views(imgs_l, s["label"], s["frame_count"])
        return clips, label, s["vid"]

    def __getitem__(self, index):
        """-> (clips [N,3,T,S,S] fp32 on `device`, label, vid), as the reference's (charades_fine.py:196); with
        host_items=True the host half only (a dict; finish it with device_collate in the main process)."""
        s = self.host_item(index)
        return s if self.host_items else self.finish(s)

    def device_collate(self, batch):
        """Main-process half of a host_items=True loader: `for samples in loader: batch = dataset.device_collate(samples)`."""
        return mt_collate_fn([self.finish(s) for s in batch])

    def __getstate__(self):                                               # the nvJPEG decoder does not travel to worker processes
        d = dict(self.__dict__)
        d["_decoder"] = None
        return d


def host_collate(batch):
    """collate_fn for host_items=True datasets: DataLoader workers run the collate function too, so it only gathers the
    host-side sample dicts; the main process turns them into a batch with `dataset.device_collate(samples)`."""
    return list(batch)


def mt_collate_fn(batch):
    """charades_fine.py:205-229: zero-pad clips [N,3,T,S,S] along T and labels [C,TL] along TL to the longest of the batch,
    mask = 1 over the valid label frames.  -> [clips [B,N,3,Tmax,S,S], labels [B,C,TLmax], masks [B,TLmax], vids]."""
    max_len_clips = max(b[0].shape[2] for b in batch)
    max_len_labels = max(b[1].shape[1] for b in batch)
    dev = batch[0][0].device
    clips = torch.zeros((len(batch),) + tuple(batch[0][0].shape[:2]) + (max_len_clips,) + tuple(batch[0][0].shape[3:]),
                        dtype=torch.float32, device=dev)
    labels = torch.zeros(len(batch), batch[0][1].shape[0], max_len_labels, dtype=torch.float32)
    masks = torch.zeros(len(batch), max_len_labels, dtype=torch.float32)
    for i, b in enumerate(batch):
        clips[i, :, :, :b[0].shape[2]] = b[0]
        labels[i, :, :b[1].shape[1]] = b[1]
        masks[i, :b[1].shape[1]] = 1
    return [clips, labels, masks, tuple(b[2] for b in batch)]
