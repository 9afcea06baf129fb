"""TEST HARNESS: the collectives of youtokentome_b200.distributed.train_distributed between host THREADS of one process
(one thread = one rank).  Only meaningful with the SIMT-emulated library, whose "device" memory is host memory: the
ranks' exchange buffers are then plain pointers into the same address space and the per-merge peer stores of the
merge loop are ordinary stores between the threads that run the ranks' kernels."""
import ctypes as C
import threading

import numpy as np


class Shared:
    def __init__(self, world):
        self.world = world
        self.slots = [None] * world
        self.barrier = threading.Barrier(world)


class ThreadComm:
    def __init__(self, shared, rank):
        self.sh, self.rank, self.world = shared, rank, shared.world

    def _exchange(self, value):
        self.sh.slots[self.rank] = value
        self.sh.barrier.wait()
        out = list(self.sh.slots)
        self.sh.barrier.wait()
        return out

    def all_gather_bytes(self, b):
        return b"".join(self._exchange(bytes(b)))

    def allreduce_sum_u64(self, ptr, n):
        arr = np.ctypeslib.as_array((C.c_uint64 * n).from_address(ptr))
        total = sum(self._exchange(arr.copy()))
        arr[:] = total

    def agree(self, ok):
        return all(self._exchange(bool(ok)))

    def all_to_all(self, ptr, counts, itemsize):
        counts = [int(c) for c in counts]
        total = sum(counts) * itemsize
        buf = np.ctypeslib.as_array((C.c_uint8 * max(total, 1)).from_address(ptr))[:total].copy() if total else np.zeros(0, np.uint8)
        cuts = np.cumsum([0] + [c * itemsize for c in counts])
        pieces = [buf[cuts[d]:cuts[d + 1]] for d in range(self.world)]
        allp = self._exchange(pieces)
        recv = [allp[src][self.rank] for src in range(self.world)]
        out = np.ascontiguousarray(np.concatenate(recv)) if recv else np.zeros(0, np.uint8)
        if out.size == 0:
            out = np.zeros(8, np.uint8)
        return out, out.ctypes.data, [len(r) // itemsize for r in recv]

    def barrier(self):
        self.sh.barrier.wait()


def run_ranks(world, fn):
    """fn(comm) on `world` threads; returns the list of results, re-raises the first exception."""
    sh = Shared(world)
    res, err = [None] * world, [None] * world

    def body(r):
        try:
            res[r] = fn(ThreadComm(sh, r))
        except BaseException as e:  # noqa: BLE001 - reported by the caller
            err[r] = e
            sh.barrier.abort()

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return res
