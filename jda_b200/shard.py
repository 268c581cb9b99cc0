"""Multi-GPU plumbing of the detect path: frames shard across ranks in contiguous
blocks, the model is replicated, and the only exchange is one all-gather of fixed-stride detection
records per batch so that every rank ends with the identical job-wide detection table.

The reference has no distributed code; the record carries exactly what its jdaResult holds per
face (c/jda.h:18-24): bbox (x, y, size), score, 2*L landmark floats -- plus the global frame id.

Works on any torch.distributed backend: NCCL with CUDA tensors on the B200 box (NVLink/NVSwitch),
gloo with CPU tensors in the CPU tests.
"""
import numpy as np

HEADER = 5  # frame, x, y, size, score


def shard_range(n_frames, rank, world):
    """contiguous block of frames owned by `rank`: frame i -> rank floor(i * world / n)."""
    lo = (n_frames * rank + world - 1) // world
    hi = (n_frames * (rank + 1) + world - 1) // world
    return lo, hi


def pack_records(results, frame0=0, landmark_n=27):
    """list of (boxes, scores, shapes) per local frame -> [n, 5 + 2L] float32 records.
    Integers up to 2^24 (frame ids, pixel coordinates) are exact in float32."""
    D = 2 * landmark_n
    counts = np.fromiter((len(r[1]) for r in results), np.int64, len(results))
    n = int(counts.sum())
    rec = np.zeros((n, HEADER + D), np.float32)
    if n:
        hit = [r for r, k in zip(results, counts) if k]
        rec[:, 0] = np.repeat(np.arange(len(results)) + frame0, counts)
        rec[:, 1:4] = np.concatenate([r[0] for r in hit])
        rec[:, 4] = np.concatenate([r[1] for r in hit])
        rec[:, HEADER:] = np.concatenate([r[2] for r in hit])
    return rec


def pack_records_flat(counts, boxes, scores, shapes, frame0=0):
    """the same records from a flat batch result (api.Cascador.detect_batch(flat=True) / jdaB200DetectBatchFlat)"""
    n = len(scores)
    rec = np.empty((n, HEADER + shapes.shape[1]), np.float32)
    rec[:, 0] = np.repeat(np.arange(len(counts)) + frame0, counts)
    rec[:, 1:4] = boxes
    rec[:, 4] = scores
    rec[:, HEADER:] = shapes
    return rec


def unpack_records(rec):
    """records -> (frame ids, boxes i32, scores, shapes)."""
    return (rec[:, 0].astype(np.int64), rec[:, 1:4].astype(np.int32), rec[:, 4].copy(),
            rec[:, HEADER:].copy())


def all_gather_records(rec, device=None, group=None):
    """One count all-gather + one padded record all-gather.  Returns the job-wide table ordered by
    (rank, local order) = global frame order for contiguous sharding.  Collective: every rank calls."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    cnt = torch.tensor([rec.shape[0]], dtype=torch.int64, device=dev)
    cnts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    cnts_h = cnts.cpu().numpy()
    mx = max(int(cnts_h.max()), 1)
    width = rec.shape[1]
    mine = torch.zeros((mx, width), dtype=torch.float32, device=dev)
    if rec.shape[0]:
        mine[:rec.shape[0]] = torch.from_numpy(rec).to(dev)
    allr = torch.empty((world * mx, width), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(allr, mine, group=group)
    allr = allr.cpu().numpy().reshape(world, mx, width)
    return np.concatenate([allr[r, :int(cnts_h[r])] for r in range(world)], axis=0)


class RecordGather:
    """The exchange step as ONE collective per batch that can run under the next batch's scan.

    Every rank sends a fixed-capacity block [cap + 1, width]: row 0 carries its record count, rows 1.. the records.
    start() uploads the block and launches the all-gather asynchronously (NCCL runs it on its own stream while the
    caller goes on to scan the next batch); finish() waits, reads the job-wide table back and returns it ordered by
    (rank, local order) = global frame order.  A batch with more records than `cap` on some rank is rare (the
    capacity doubles past the largest count seen): finish() then repeats the exchange synchronously with a larger
    block, so the result never depends on the capacity."""

    def __init__(self, width, device=None, group=None, cap=256):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.width, self.group, self.cap = width, group, cap
        self.dev = torch.device(device) if device is not None else torch.device("cpu")
        self.world = dist.get_world_size(group)
        self.pending = None
        self.exchanges = 0
        self._host = {}
        self._back = {}
        # CUDA: the exchange lives on its own stream (upload -> all-gather -> read-back into pinned memory -> event),
        # so neither start() nor finish() issues a blocking copy: finish() waits on the event only
        self._side = torch.cuda.Stream(self.dev) if self.dev.type == "cuda" else None

    def _launch(self, rec, cap, async_op):
        torch = self.torch
        host = self._host.get(cap)
        if host is None:      # one pinned staging block per capacity, reused by every batch
            host = torch.zeros((cap + 1, self.width), dtype=torch.float32)
            if self.dev.type == "cuda":
                host = host.pin_memory()
            self._host[cap] = host
        n = min(rec.shape[0], cap)
        host[0, 0] = float(rec.shape[0])     # exact up to 2^24 records per rank
        if n:
            host[1:n + 1] = torch.from_numpy(np.ascontiguousarray(rec[:n], np.float32))
        self.exchanges += 1
        if self._side is None:               # CPU tensors (gloo)
            recv = torch.empty((self.world * (cap + 1), self.width), dtype=torch.float32)
            work = self.dist.all_gather_into_tensor(recv, host.clone(), group=self.group, async_op=async_op)
            return dict(rec=rec, cap=cap, recv=recv, work=work, event=None, back=None)
        back = self._back.get(cap)
        if back is None:
            back = torch.empty((self.world * (cap + 1), self.width), dtype=torch.float32).pin_memory()
            self._back[cap] = back
        with torch.cuda.stream(self._side):
            send = host.to(self.dev, non_blocking=True)
            recv = torch.empty((self.world * (cap + 1), self.width), dtype=torch.float32, device=self.dev)
            work = self.dist.all_gather_into_tensor(recv, send, group=self.group, async_op=True)
            work.wait()                      # stream-level: the side stream waits for NCCL, the host does not
            back.copy_(recv, non_blocking=True)
            event = torch.cuda.Event()
            event.record(self._side)
        return dict(rec=rec, cap=cap, recv=recv, send=send, work=None, event=event, back=back)

    @staticmethod
    def _next_cap(largest):
        """Block capacity after seeing a rank send `largest` records.  Detections (a few hundred per batch): twice the
        count, a power of two.  Mining (tens of thousands per batch, and every padded row crosses NVLink and PCIe to
        every rank): a quarter above the count -- batches of one job differ by a few per cent, and a batch that does
        overflow is simply exchanged again."""
        if largest < 4096:
            return 1 << int(np.ceil(np.log2(max(largest * 2, 1))))
        return (int(largest * 1.25) + 4095) // 4096 * 4096

    def start(self, rec):
        assert self.pending is None, "finish() the previous exchange first"
        self.pending = self._launch(rec, self.cap, True)

    def finish(self):
        p, self.pending = self.pending, None
        if p is None:
            return None
        while True:
            if p["work"] is not None:
                p["work"].wait()
            if p["event"] is not None:
                p["event"].synchronize()
                # (a view of the pinned read-back block: np.concatenate below copies the used rows out, and the block is
                # not written again before the next start())
                allr = p["back"].numpy().reshape(self.world, p["cap"] + 1, self.width)
            else:
                allr = p["recv"].numpy().reshape(self.world, p["cap"] + 1, self.width)
            counts = allr[:, 0, 0].astype(np.int64)
            if counts.max() <= p["cap"]:
                self.cap = max(self.cap, self._next_cap(int(counts.max())))
                return np.concatenate([allr[r, 1:1 + int(counts[r])] for r in range(self.world)], axis=0)
            # some rank overflowed its block: every rank sees the same counts, so every rank repeats the exchange
            cap = self._next_cap(int(counts.max()))
            self.cap = max(self.cap, cap)
            p = self._launch(p["rec"], cap, False)
