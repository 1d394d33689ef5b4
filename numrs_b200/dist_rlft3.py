"""Slab-decomposed rlft3 (and 3-D complex fourn, `kind="fourn"`) across the GPUs of one box: one process per GPU
(torch.distributed for the plumbing: IPC handle exchange, barriers; NCCL only in mode "nccl").

Forward (isign=+1): rank r passes its nn2-slab data[:, r*nn2/G:(r+1)*nn2/G, :] (contiguous
[nn1][nn2/G][nn3] real) and gets back, in the same buffer, its nn1-slab of the spectrum
([nn1/G][nn2][nn3/2] complex = rows r*nn1/G.. of the reference's layout) plus speq [nn1/G][2*nn2].
Inverse (isign=-1) is the mirror image.  Exactly one exchange per direction:

  mode "fused" (default): the last FFT pass of stage 0 stores its output straight into the owning
      peer's receive buffer through NVLink (CUDA IPC mapped peer memory) from the kernel epilogue,
      so the transfer overlaps the transform tile by tile.  With `pull=True` (default) only the low-z half of every
      block is pushed that way; the other half is written to the rank's own send buffer and PULLED by the consumer's
      stage-1 pass with loads over NVLink, so the link carries half the bytes during each of the two passes of a
      direction instead of all of them in front of the second one.  The barrier between stage 0 and stage 1
      is a pair of one-warp kernels exchanging epoch flags through the same peer mappings
      (`barrier="flags"`, default: no collective at all on the data path) or a stream-ordered
      1-element NCCL all-reduce (`barrier="nccl"`).
      `chunks` > 1 pipelines the exchange: the volume is cut into z-ranges, stage 1 of chunk c (HBM-bound, local)
      runs on a second stream under the NVLink-bound stage-0 stores of chunk c + 1; every chunk has its own
      epoch flags.
  mode "dma": stage 0 writes a chunk-major send buffer, copy engines (not SMs) push every (peer, chunk) piece into
      the peer's receive buffer over NVLink, a flag per chunk follows the copies, and stage 1 of the chunk runs on
      a high-priority side stream -- the local passes overlap the transfer (nrb_slab_exec_dma does the whole
      direction in one C call to keep the launch overhead down).
  mode "nccl": stage 0 writes a send buffer, torch.distributed.all_to_all_single moves the blocks.
"""
import torch
import torch.distributed as dist


class SlabRlft3:
    def __init__(self, lib, nn1, nn2, nn3, mode="fused", barrier="flags", chunks=1, kind="rlft3", pull=True):
        self.lib = lib
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.dims = (nn1, nn2, nn3)
        self.kind = kind
        self.plan = lib.slab_create(nn1, nn2, nn3, self.world, self.rank, kind=kind)
        self.local_doubles = self.plan.local_doubles()
        self.speq_doubles = self.plan.speq_doubles()
        self.xchg_doubles = self.plan.xchg_doubles()
        self.mode = mode
        self.barrier = barrier
        self._flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._call = 0
        self.chunks = chunks if (mode in ("fused", "dma") and barrier == "flags" and self.world > 1) else 1
        if mode == "dma":
            self.plan.set_dma(self.chunks)
        elif self.chunks > 1:
            self.plan.set_chunks(self.chunks)
            self._side = torch.cuda.Stream(priority=-1)       # stage 1 pieces: short, HBM-bound -> scheduled first
            self._ev_go, self._ev_done = torch.cuda.Event(), torch.cuda.Event()
        self.pull = bool(pull) and mode == "fused" and self.chunks == 1 and self.world > 1
        if mode in ("fused", "dma"):
            # two receive buffers, alternated per call, so a peer still reading call k's data in its
            # stage 1 is never overwritten by call k+1's stage 0 (ordered by call k+1's barrier)
            self._own, self._peers = [], []
            self._own_send, self._send_peers = [], []

            def shared(nbytes):
                own = lib.device_alloc(nbytes)                     # zeroed: the flag array starts at epoch 0
                handles = [None] * self.world
                dist.all_gather_object(handles, lib.ipc_export(own))
                return own, [own if r == self.rank else lib.ipc_import(h) for r, h in enumerate(handles)]
            for _ in range(2):
                own, peers = shared(self.plan.recv_bytes())
                self._own.append(own)
                self._peers.append(peers)
                if self.pull:      # push + pull: the send buffers are read by the peers' stage 1, double buffered the same way
                    own, peers = shared(8 * self.xchg_doubles)
                    self._own_send.append(own)
                    self._send_peers.append(peers)
            dist.barrier()
        elif mode == "nccl":
            self.send = torch.empty(self.xchg_doubles, dtype=torch.float64, device="cuda")
            self.recv = torch.empty(self.xchg_doubles, dtype=torch.float64, device="cuda")
        else:
            raise ValueError(mode)

    def a2a_bytes_per_gpu(self):
        return 8.0 * self.xchg_doubles * (self.world - 1) / self.world

    def transform(self, slab, speq, isign):
        """slab, speq: torch float64 CUDA tensors (local_doubles / speq_doubles; speq = None for kind "fourn"); in
        place; enqueues on the current stream."""
        st = torch.cuda.current_stream().cuda_stream
        speq_ptr = speq.data_ptr() if speq is not None else 0
        if self.mode == "dma":
            self.plan.set_peers(self._peers[self._call & 1])
            self._call += 1
            self.plan.exec_dma(isign, slab.data_ptr(), speq_ptr, (self._call + 1) // 2, st)
        elif self.mode == "fused":
            peers = self._peers[self._call & 1]
            if self.pull:
                self.plan.set_send_peers(self._send_peers[self._call & 1])
            self._call += 1
            self.plan.set_peers(peers)
            if self.chunks > 1:
                self._pipelined(slab, speq, isign, (self._call + 1) // 2)
                return
            if self.barrier == "flags":
                # stage 0 (stores land in the peers' receive buffers), epoch-flag barrier, stage 1: one C call
                self.plan.exec(isign, slab.data_ptr(), speq_ptr, (self._call + 1) // 2, st)   # epoch per receive buffer: 1, 2, ...
                return
            self.plan.stage(0, isign, slab.data_ptr(), speq_ptr, 0, 0, st)
            dist.all_reduce(self._flag)                  # stream-ordered NCCL barrier
            self.plan.stage(1, isign, slab.data_ptr(), speq_ptr, 0, 0, st)
        else:
            self.plan.stage(0, isign, slab.data_ptr(), speq_ptr, self.send.data_ptr(), 0, st)
            dist.all_to_all_single(self.recv, self.send)
            self.plan.stage(1, isign, slab.data_ptr(), speq_ptr, 0, self.recv.data_ptr(), st)

    def _pipelined(self, slab, speq, isign, epoch):
        main = torch.cuda.current_stream()
        st, sd = main.cuda_stream, self._side.cuda_stream
        d, q, C = slab.data_ptr(), (speq.data_ptr() if speq is not None else 0), self.chunks
        self.plan.stage_part(0, -1, isign, d, q, st)              # forward: z pass (slab -> work)
        self._ev_go.record(main)
        self._side.wait_event(self._ev_go)
        for c in range(C):
            self.plan.stage_part(0, c, isign, d, q, st)           # x (or y) pass of chunk c, stores go to the peers
            self.plan.barrier_chunk(0, c, epoch, st)              # chunk c of this rank has landed everywhere
            self.plan.barrier_chunk(1, c, epoch, sd)              # side stream: chunk c of every rank has landed here
            self.plan.stage_part(1, c, isign, d, q, sd)
        self._ev_done.record(self._side)
        main.wait_event(self._ev_done)
        self.plan.stage_part(0, C, isign, d, q, st)               # inverse: z pass (work -> slab)

    def close(self):
        torch.cuda.synchronize()
        dist.barrier()
        if self.mode in ("fused", "dma"):
            for peers in self._peers + self._send_peers:
                for r, p in enumerate(peers):
                    if r != self.rank:
                        self.lib.ipc_release(p)
            dist.barrier()
            for own in self._own + self._own_send:
                self.lib.device_free(own)
        self.plan.destroy()
