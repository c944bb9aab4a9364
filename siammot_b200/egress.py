"""Result egress for a whole clip.

The reference turns every frame's BoxList into Python objects one element at a time -- ``o.resize([w, h]).convert('xywh')``,
``.to(cpu)``, then ``boxlists_to_entities`` with three ``.item()`` / ``.tolist()`` calls per box
(siammot/engine/inferencer.py:64-70, siammot/utils/boxlists_to_entities.py:22-35) -- and, per video, filters short / unconfident
tracks with a Python loop over ids and entities (``DatasetInference._postprocess_tracks``, inferencer.py:134-153).  At hundreds
of frames per second that is the dominant host cost after the engine itself.  Here a clip's results become flat numpy columns
in one pass (same IEEE operations as BoxList.resize / convert, so the numbers are identical), the track filter is a
vectorised group-by, and entity objects are only materialised on request.

    tracks = clip_to_tracks(model.forward_clip(frames), video_width, video_height)      # model.results_on_host = True
    kept = postprocess_tracks(tracks)                                                   # len >= 5 and mean confidence >= 0.7
    entities = to_entities(kept)                                                        # reference-style objects, if needed
"""
import numpy as np

TO_REMOVE = 1.0


class ClipTracks(object):
    """Columns over all boxes of a clip, frame-major: frame_num i64, time f64, id i64, label i64, confidence f32,
    bbox (n,4) f32 in xywh (legacy +1 widths) at the original video size."""
    __slots__ = ("frame_num", "time", "id", "label", "confidence", "bbox")

    def __init__(self, frame_num, time, id, label, confidence, bbox):
        self.frame_num, self.time, self.id, self.label, self.confidence, self.bbox = frame_num, time, id, label, confidence, bbox

    def __len__(self):
        return int(self.id.shape[0])

    def select(self, mask_or_index):
        return ClipTracks(*(getattr(self, k)[mask_or_index] for k in self.__slots__))


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def clip_to_tracks(results, width, height, first_frame_idx=0, timestamps=None):
    """results: the per-frame BoxLists of one clip (xyxy at the network input size, fields scores / ids / labels), ideally CPU
    ones (``model.results_on_host = True``).  Boxes are rescaled to (width, height) and converted to xywh exactly as
    ``BoxList.resize`` + ``convert('xywh')`` do (inferencer.py:65)."""
    n_frames = len(results)
    if timestamps is None:
        timestamps = list(range(first_frame_idx, first_frame_idx + n_frames))
    counts = np.fromiter((len(r) for r in results), dtype=np.int64, count=n_frames)
    total = int(counts.sum())
    frame_num = np.repeat(np.arange(first_frame_idx, first_frame_idx + n_frames, dtype=np.int64), counts)
    time = np.repeat(np.asarray(timestamps, dtype=np.float64), counts)
    bbox = np.empty((total, 4), dtype=np.float32)
    conf = np.empty((total,), dtype=np.float32)
    ids = np.full((total,), -1, dtype=np.int64)
    labels = np.empty((total,), dtype=np.int64)
    o = 0
    for r, c in zip(results, counts.tolist()):
        if c == 0:
            continue
        assert r.mode == "xyxy"
        b = _np(r.bbox).astype(np.float32, copy=False)
        rw, rh = float(width) / float(r.size[0]), float(height) / float(r.size[1])
        if rw == rh:                                         # BoxList.resize: one ratio for all four coordinates
            s = b * np.float32(rw)
        else:                                                # ... or per axis
            s = b * np.array([rw, rh, rw, rh], dtype=np.float32)
        out = bbox[o:o + c]
        out[:, 0:2] = s[:, 0:2]
        out[:, 2] = s[:, 2] - s[:, 0] + np.float32(TO_REMOVE)   # convert('xywh')
        out[:, 3] = s[:, 3] - s[:, 1] + np.float32(TO_REMOVE)
        conf[o:o + c] = _np(r.get_field("scores"))
        labels[o:o + c] = _np(r.get_field("labels"))
        if r.has_field("ids"):
            ids[o:o + c] = _np(r.get_field("ids"))
        o += c
    return ClipTracks(frame_num, time, ids, labels, conf, bbox)


def postprocess_tracks(tracks, track_len=5, track_conf=0.7):
    """DatasetInference._postprocess_tracks (inferencer.py:134-153): keep the entities of ids >= 0 that appear at least
    ``track_len`` times with a mean confidence >= ``track_conf``; untracked boxes (id -1) are dropped.  Returned grouped by id
    (ascending), frames ascending within an id."""
    m = tracks.id >= 0
    if not m.any():
        return tracks.select(m)
    idx = np.nonzero(m)[0]
    uniq, inv, cnt = np.unique(tracks.id[idx], return_inverse=True, return_counts=True)
    mean = np.bincount(inv, weights=tracks.confidence[idx].astype(np.float64), minlength=uniq.shape[0]) / cnt
    good = (cnt >= track_len) & (mean >= track_conf)
    keep = idx[good[inv]]
    order = np.lexsort((tracks.frame_num[keep], tracks.id[keep]))
    return tracks.select(keep[order])


class Entity(object):
    """The members of gluoncv's AnnoEntity that boxlists_to_entities fills (boxlists_to_entities.py:24-34)."""
    __slots__ = ("bbox", "confidence", "labels", "id", "frame_num", "time")


def to_entities(tracks, class_table=None, entity_cls=Entity):
    """Reference-style entity objects (one column-wise ``tolist()`` instead of three ``.item()`` per box)."""
    if class_table is None:
        class_table = ["person"]                             # boxlists_to_entities.py:15-17
    out = []
    for bb, cf, lb, i, fn, tm in zip(tracks.bbox.tolist(), tracks.confidence.tolist(), tracks.label.tolist(), tracks.id.tolist(),
                                     tracks.frame_num.tolist(), tracks.time.tolist()):
        e = entity_cls()
        e.bbox, e.confidence, e.labels, e.id, e.frame_num, e.time = bb, cf, {class_table[lb - 1]: cf}, i, fn, tm
        out.append(e)
    return out
