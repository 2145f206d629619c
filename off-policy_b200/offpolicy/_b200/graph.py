"""Whole-step CUDA graph: [sample ->] QMIX step [-> PER write-back] [-> soft update] captured once, replayed per step.

`mx_graph_capture` records the library's own launch sequence on a dedicated (non-default) stream; after that a learner
step costs one `cudaGraphLaunch` and no host work at all (indices come from the device-resident MT19937 stream).
"""
import ctypes as C

import torch

from offpolicy._b200 import capi

SAMPLE_UNIFORM, SAMPLE_PER, SOFT_UPDATE, PER_WRITEBACK = 1, 2, 4, 8


class StepGraph(object):
    def __init__(self, buffer, trainer, batch_size, beta=0.4, soft_update=True, p_id="policy_0"):
        lib = capi.lib()
        pb = buffer.policy_buffers[p_id]
        per = bool(getattr(trainer, "use_per", False))
        self.flags = (SAMPLE_PER | PER_WRITEBACK if per else SAMPLE_UNIFORM) | (SOFT_UPDATE if soft_update else 0)
        self.stream = torch.cuda.Stream(device=capi.device())
        self.stream.wait_stream(torch.cuda.current_stream(capi.device()))
        g = C.c_void_p()
        capi.check(lib.mx_graph_capture(pb.handle, trainer.handle, int(batch_size), float(beta), self.flags,
                                        C.c_void_p(self.stream.cuda_stream), C.byref(g)))
        self.handle = g
        self.num_kernels = int(lib.mx_graph_num_kernels(g))
        self._keep = (buffer, trainer)

    def launch(self):
        capi.check(capi.lib().mx_graph_launch(self.handle, C.c_void_p(self.stream.cuda_stream)))

    def synchronize(self):
        self.stream.synchronize()

    def close(self):
        if self.handle:
            capi.lib().mx_graph_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
