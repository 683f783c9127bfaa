"""Peer-memory communicator for the rephrase-sharded decision (cvb_comm_* / cvb_allgather_select, include/coverb200.h):
torch.distributed is only used ONCE, to exchange the CUDA IPC handles of the per-rank mailboxes; every decision after
that is one kernel per rank that stores its score / action slice into its peers' HBM over NVLink and selects."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .cover import rephrase_shard


class PeerGather:
    def __init__(self, max_slot_floats: int, device=None, group=None):
        import torch.distributed as dist
        self.lib = _lib.load()
        L = self.lib
        L.cvb_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.cvb_comm_local_handle.argtypes = [C.c_void_p, C.c_void_p]
        L.cvb_comm_open_peers.argtypes = [C.c_void_p, C.c_void_p]
        L.cvb_comm_destroy.argtypes = [C.c_void_p]
        L.cvb_comm_destroy.restype = None
        L.cvb_allgather_select.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self._c = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.cvb_comm_create(self.rank, self.world, int(max_slot_floats), C.byref(self._c)))
            nb = L.cvb_comm_handle_bytes()
            mine = (C.c_ubyte * nb)()
            _lib.check(L.cvb_comm_local_handle(self._c, mine))
            if self.world > 1:
                local = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=self.device)
                allh = torch.empty(self.world * nb, dtype=torch.uint8, device=self.device)
                dist.all_gather_into_tensor(allh, local, group=group)
                buf = (C.c_ubyte * (self.world * nb)).from_buffer_copy(bytes(allh.cpu().tolist()))
                _lib.check(L.cvb_comm_open_peers(self._c, buf))
                dist.barrier(group=group)  # every mailbox is mapped everywhere before the first push
            else:
                _lib.check(L.cvb_comm_open_peers(self._c, None))

    def __call__(self, local_scores, local_actions, R: int, K: int):
        """local_scores f32 [n_loc]; local_actions f32 [n_loc, ...] or None.  Returns (scores [R*K], actions [R*K, ...] or
        None, group_mean [R], best_idx i32 [1], best_score [1]) - the same values on every rank.  Asynchronous."""
        a, b = rephrase_shard(R, self.world, self.rank)
        n_loc = (b - a) * K
        assert local_scores.dtype == torch.float32 and local_scores.numel() == n_loc and local_scores.is_cuda
        dev = local_scores.device
        scores = torch.empty(R * K, dtype=torch.float32, device=dev)
        actions, act_floats = None, 0
        if local_actions is not None:
            local_actions = local_actions.contiguous()
            assert local_actions.dtype == torch.float32 and local_actions.shape[0] == n_loc
            act_floats = local_actions[0].numel()
            actions = torch.empty((R * K, *local_actions.shape[1:]), dtype=torch.float32, device=dev)
        gmean = torch.empty(R, dtype=torch.float32, device=dev)
        bidx = torch.zeros(1, dtype=torch.int32, device=dev)
        bscore = torch.zeros(1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.cvb_allgather_select(self._c, _lib.ptr(local_scores.contiguous()), _lib.ptr(local_actions),
                                                     act_floats, R, K, _lib.ptr(scores), _lib.ptr(actions), _lib.ptr(gmean),
                                                     _lib.ptr(bidx), _lib.ptr(bscore), _lib.stream_ptr()))
        return scores, actions, gmean, bidx, bscore

    def close(self):
        if self._c.value:
            self.lib.cvb_comm_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
