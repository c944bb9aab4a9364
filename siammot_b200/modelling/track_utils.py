"""TrackUtils / TrackPool: host-side track bookkeeping with the reference's interface
(/root/reference/siammot/modelling/track_head/track_utils.py:12-269).

The id allocator, the active set, the dormant table and the per-id cache decide which integer id
every box gets, so their container semantics (a ``set`` of active ids, an insertion-ordered ``dict``
of dormant ids, iteration over ``set(dict.keys())``) are kept identical to the reference: the order
of dormant tracks in the memory -- and through NMS tie-breaking the output -- depends on them."""
import torch

from ..structures import BoxList


class TrackUtils(object):
    def __init__(self, search_expansion=1.0, min_search_wh=128, pad_pixels=256):
        self.search_expansion = search_expansion
        self.min_search_wh = min_search_wh
        self.pad_pixels = pad_pixels

    def search_region(self, boxes):
        """boxes (N,4) fp32 CPU xyxy (image frame) -> SR boxes in the padded frame.
        update_boxes_in_pad_images + extend_bbox (track_utils.py:62-85,109-135), same op order."""
        sr = boxes + float(self.pad_pixels)
        w = sr[:, 2] - sr[:, 0] + 1
        h = sr[:, 3] - sr[:, 1] + 1
        w_ext = torch.max((self.min_search_wh - w) / (self.search_expansion * 2.), w * (self.search_expansion / 2.))
        h_ext = torch.max((self.min_search_wh - h) / (self.search_expansion * 2.), h * (self.search_expansion / 2.))
        return torch.stack((sr[:, 0] - w_ext, sr[:, 1] - h_ext, sr[:, 2] + w_ext, sr[:, 3] + h_ext), dim=1)

    def search_region_np(self, boxes):
        """numpy fp32 twin of search_region (identical IEEE operations, lower host overhead)."""
        import numpy as np
        sr = boxes + np.float32(self.pad_pixels)
        w = sr[:, 2] - sr[:, 0] + np.float32(1)
        h = sr[:, 3] - sr[:, 1] + np.float32(1)
        se = float(self.search_expansion)
        w_ext = np.maximum((np.float32(self.min_search_wh) - w) / np.float32(se * 2.), w * np.float32(se / 2.))
        h_ext = np.maximum((np.float32(self.min_search_wh) - h) / np.float32(se * 2.), h * np.float32(se / 2.))
        return np.stack((sr[:, 0] - w_ext, sr[:, 1] - h_ext, sr[:, 2] + w_ext, sr[:, 3] + h_ext), axis=1).astype(np.float32)

    def update_boxes_in_pad_images(self, boxlists):
        out = []
        for bl in boxlists:
            assert bl.mode == "xyxy"
            w, h = bl.size
            nb = BoxList(bl.bbox + float(self.pad_pixels), [int(w + 2 * self.pad_pixels), int(h + 2 * self.pad_pixels)], "xyxy")
            for f in bl.fields():
                nb.add_field(f, bl.get_field(f))
            out.append(nb)
        return out

    def extend_bbox(self, in_box):
        for bl in in_box:
            b = bl.bbox
            w = b[:, 2] - b[:, 0] + 1
            h = b[:, 3] - b[:, 1] + 1
            w_ext = torch.max((self.min_search_wh - w) / (self.search_expansion * 2.), w * (self.search_expansion / 2.))
            h_ext = torch.max((self.min_search_wh - h) / (self.search_expansion * 2.), h * (self.search_expansion / 2.))
            b[:, 0] -= w_ext
            b[:, 1] -= h_ext
            b[:, 2] += w_ext
            b[:, 3] += h_ext
        return in_box

    def pad_feature(self, f):
        raise RuntimeError("pad_feature is eliminated: libsmot's ROIAlign samples the zero padding virtually "
                           "(smot_roi_align pad[]); see INTEGRATION.md")


class TrackPool(object):
    def __init__(self, active_ids=None, max_entangle_length=10, max_dormant_frames=1):
        self._max_dormant_frames = max_dormant_frames
        self._max_entangle_length = max_entangle_length
        self.reset()

    def reset(self):
        self._active_ids = set()
        self._kill_ids = set()
        self._dormant_ids = {}
        self._embedding = None
        self._cache = {}
        self._max_id = -1
        self._frame_idx = 0

    def suspend_track(self, track_id):
        if track_id not in self._active_ids:
            raise ValueError
        self._active_ids.remove(track_id)
        self._dormant_ids[track_id] = self._frame_idx - 1

    def expire_tracks(self):
        for track_id, last_active in list(self._dormant_ids.items()):
            if self._frame_idx - last_active >= self._max_dormant_frames:
                self._dormant_ids.pop(track_id)
                self._kill_ids.add(track_id)
                self._cache.pop(track_id, None)

    def increment_frame(self, value=1):
        self._frame_idx += value

    def update_cache(self, cache):
        """cache: dict id -> per-track state.  (The reference takes the memory tuple and splits it
        per id, track_utils.py:180-197; the engine passes the already split rows.)"""
        for track_id, entry in cache.items():
            self._cache[track_id] = entry

    def resume_track(self, track_id):
        if track_id not in self._dormant_ids or track_id in self._active_ids:
            raise ValueError
        self._active_ids.add(track_id)
        self._dormant_ids.pop(track_id)

    def kill_track(self, track_id):
        if track_id not in self._active_ids:
            raise ValueError
        self._active_ids.remove(track_id)
        self._kill_ids.add(track_id)
        self._cache.pop(track_id, None)

    def start_track(self):
        new_id = self._max_id + 1
        self._max_id = new_id
        self._active_ids.add(new_id)
        return new_id

    def get_active_ids(self):
        return self._active_ids

    def get_dormant_ids(self):
        return set(self._dormant_ids.keys())

    def get_cache(self):
        return self._cache

    def activate_tracks(self, track_id):
        if track_id in self._active_ids or track_id not in self._dormant_ids:
            raise ValueError
        self._active_ids.add(track_id)
        self._dormant_ids.pop(track_id)


def build_track_utils(cfg):
    T = cfg.MODEL.TRACK_HEAD
    track_utils = TrackUtils(search_expansion=T.SEARCH_REGION - 1., min_search_wh=T.MINIMUM_SREACH_REGION,
                             pad_pixels=T.PAD_PIXELS)
    track_pool = TrackPool(max_dormant_frames=T.MAX_DORMANT_FRAMES)
    return track_utils, track_pool
