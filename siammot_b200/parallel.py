"""Multi-GPU: one process per GPU, each owning whole video streams (weights replicated, no data-path
collective -- the reference has no multi-GPU inference at all, inferencer.py:156 is a todo).  The
single exchange is a per-clip all-gather of fixed-size track-state records:

    record = [x1, y1, x2, y2, score, id, label, valid]  (8 x fp32; ids < 2^24 are exact in fp32)

NCCL over NVLink when the tensors are on GPUs; the same code runs on gloo/CPU tensors (tests)."""
import torch
import torch.distributed as dist

RECORD = 8


def shard_streams(num_streams, rank, world_size):
    """GPU g owns streams {g, g + world, g + 2*world, ...}."""
    return list(range(rank, num_streams, world_size))


def pack_track_states(result, max_tracks, device=None):
    """BoxList (fields scores, ids, labels) -> (max_tracks, 8) fp32 record array, tracked boxes only."""
    device = device or result.bbox.device
    rec = torch.zeros((max_tracks, RECORD), dtype=torch.float32, device=device)
    ids = result.get_field("ids")
    sel = (ids >= 0).nonzero().squeeze(1)[:max_tracks]
    n = sel.numel()
    if n:
        rec[:n, 0:4] = result.bbox[sel].to(device)
        rec[:n, 4] = result.get_field("scores")[sel].to(device)
        rec[:n, 5] = ids[sel].to(device, torch.float32)
        rec[:n, 6] = result.get_field("labels")[sel].to(device, torch.float32)
        rec[:n, 7] = 1.0
    return rec


def exchange_device(result=None, group=None):
    """The device the records must live on for the process group's backend: NCCL moves device memory only (a CPU tensor
    handed to it fails with "no backend for device cpu" -- results_on_host BoxLists are CPU tensors), gloo moves host memory."""
    if dist.is_available() and dist.is_initialized():
        backend = str(dist.get_backend(group)).lower()
        if "nccl" in backend:
            return torch.device("cuda", torch.cuda.current_device())
        if "gloo" in backend:
            return torch.device("cpu")
    return result.bbox.device if result is not None else torch.device("cpu")


def gather_track_states(result, max_tracks=128, group=None):
    """All ranks receive every rank's records: returns a (world, max_tracks, 8) tensor on the exchange device
    (the rank's GPU under NCCL, the host under gloo), whatever device the BoxList itself lives on."""
    rec = pack_track_states(result, max_tracks, device=exchange_device(result, group))
    if not (dist.is_available() and dist.is_initialized()):
        return rec[None]
    world = dist.get_world_size(group)
    parts = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(parts, rec, group=group)
    return torch.stack(parts)


def unpack_track_states(records):
    """(world, max_tracks, 8) -> list over ranks of dict(boxes, scores, ids, labels)."""
    out = []
    for r in records:
        v = r[:, 7] > 0.5
        out.append(dict(boxes=r[v, 0:4], scores=r[v, 4], ids=r[v, 5].round().to(torch.int64),
                        labels=r[v, 6].round().to(torch.int64)))
    return out
