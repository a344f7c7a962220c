"""Pipeline schedules as per-stage instruction streams.

Parity target: reference ``runtime/pipe/schedule.py`` (``TrainSchedule :189`` 1F1B, ``InferenceSchedule :135``,
instruction classes ``:327-487``).  The generator here is derived from the classic formulation -- warm-up
forwards, steady one-forward-one-backward, cool-down backwards -- rather than the reference's even/odd step
arithmetic; both produce the same per-stage ordering of work and the same peak number of live activations
(``min(stages - stage_id, micro_batches)`` buffers).
"""
from abc import ABC, abstractmethod


class PipeInstruction:

    def __init__(self, **kwargs):
        self.name = self.__class__.__name__
        self.kwargs = kwargs
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __repr__(self):
        args = ", ".join(f"{k}={v!r}" for k, v in self.kwargs.items())
        return f"{self.name}({args})"

    def __eq__(self, other):
        return type(self) is type(other) and self.kwargs == other.kwargs


class OptimizerStep(PipeInstruction):
    pass


class ReduceGrads(PipeInstruction):
    pass


class ReduceTiedGrads(PipeInstruction):
    pass


class BufferOpInstruction(PipeInstruction):

    def __init__(self, buffer_id, **kwargs):
        super().__init__(buffer_id=buffer_id, **kwargs)


class LoadMicroBatch(BufferOpInstruction):
    pass


class ForwardPass(BufferOpInstruction):
    pass


class BackwardPass(BufferOpInstruction):
    pass


class SendActivation(BufferOpInstruction):
    pass


class RecvActivation(BufferOpInstruction):
    pass


class SendGrad(BufferOpInstruction):
    pass


class RecvGrad(BufferOpInstruction):
    pass


class PipeSchedule(ABC):

    def __init__(self, micro_batches, stages, stage_id):
        super().__init__()
        self.micro_batches = micro_batches
        self.stages = stages
        self.stage_id = stage_id
        self.prev_stage = stage_id - 1
        self.next_stage = stage_id + 1

    @abstractmethod
    def steps(self):
        ...

    def num_pipe_buffers(self):
        return self.micro_batches

    @property
    def stage(self):
        return self.stage_id

    @property
    def num_stages(self):
        return self.stages

    @property
    def num_micro_batches(self):
        return self.micro_batches

    @property
    def is_first_stage(self):
        return self.stage_id == 0

    @property
    def is_last_stage(self):
        return self.stage_id == self.stages - 1

    def _buffer_idx(self, micro_batch_id):
        assert 0 <= micro_batch_id < self.micro_batches
        return micro_batch_id % self.num_pipe_buffers()

    def __iter__(self):
        self.it = None
        return self

    def __next__(self):
        if self.it is None:
            self.it = self.steps()
        return next(self.it)


class InferenceSchedule(PipeSchedule):
    """Forward-only pipelining with two alternating buffers."""

    def steps(self):
        for mb in range(self.micro_batches):
            buf = mb % 2
            cmds = []
            if self.is_first_stage or self.is_last_stage:
                cmds.append(LoadMicroBatch(buf))
            if not self.is_first_stage:
                cmds.append(RecvActivation(buf))
            cmds.append(ForwardPass(buf))
            if not self.is_last_stage:
                cmds.append(SendActivation(buf))
            yield cmds

    def num_pipe_buffers(self):
        return 2


class TrainSchedule(PipeSchedule):
    """1F1B: at most ``stages - stage_id`` activations are alive on a stage."""

    def _fwd(self, mb):
        buf = self._buffer_idx(mb)
        cmds = []
        if not self.is_first_stage:
            cmds.append(RecvActivation(buf))
        if self.is_first_stage or self.is_last_stage:
            cmds.append(LoadMicroBatch(buf))
        cmds.append(ForwardPass(buf))
        if not self.is_last_stage:
            cmds.append(SendActivation(buf))
        return cmds

    def _bwd(self, mb):
        buf = self._buffer_idx(mb)
        cmds = []
        if not self.is_last_stage:
            cmds.append(RecvGrad(buf))
        cmds.append(BackwardPass(buf))
        if not self.is_first_stage:
            cmds.append(SendGrad(buf))
        return cmds

    def steps(self):
        warm = min(self.stages - 1 - self.stage_id, self.micro_batches)
        for mb in range(warm):
            yield self._fwd(mb)
        for i in range(self.micro_batches - warm):
            yield self._fwd(warm + i)
            yield self._bwd(i)
        for mb in range(self.micro_batches - warm, self.micro_batches):
            yield self._bwd(mb)
        yield [ReduceTiedGrads(), ReduceGrads(), OptimizerStep()]

    def num_pipe_buffers(self):
        return max(2, min(self.stages - self.stage_id, self.micro_batches))


class DataParallelSchedule(PipeSchedule):
    """Degenerate single-stage schedule (plain gradient accumulation)."""

    def steps(self):
        for mb in range(self.micro_batches):
            cmds = [LoadMicroBatch(0), ForwardPass(0), BackwardPass(0)]
            if mb == self.micro_batches - 1:
                cmds.extend([ReduceGrads(), OptimizerStep()])
            yield cmds

    def num_pipe_buffers(self):
        return 1


class SendActivationRecvGrad(PipeInstruction):
    """Fused exchange with the next stage (send ``send_buffer``'s activation, receive ``recv_buffer``'s grad)."""


class SendGradRecvActivation(PipeInstruction):
    """Fused exchange with the previous stage."""


def fuse_exchanges(cmds):
    """Rewrite a flat instruction stream so that a send immediately followed by a receive on the same link
    becomes ONE exchange instruction.  In 1F1B steady state neighbouring stages both want to send first
    (activation forward, gradient backward); posting send+recv together removes the circular wait."""
    out, i = [], 0
    while i < len(cmds):
        a = cmds[i]
        b = cmds[i + 1] if i + 1 < len(cmds) else None
        if isinstance(a, SendActivation) and isinstance(b, RecvGrad) and a.buffer_id != b.buffer_id:
            # (same buffer = the SAME micro batch: its gradient only exists after the next stage has consumed this very
            # activation -- a dependency, not an exchange; e.g. one micro batch per step. Those stay sequential.)
            out.append(SendActivationRecvGrad(send_buffer=a.buffer_id, recv_buffer=b.buffer_id))
            i += 2
        elif isinstance(a, SendGrad) and isinstance(b, RecvActivation):
            out.append(SendGradRecvActivation(send_buffer=a.buffer_id, recv_buffer=b.buffer_id))
            i += 2
        else:
            out.append(a)
            i += 1
    return out
