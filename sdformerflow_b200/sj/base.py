"""Step-mode / memory protocol (mirrors spikingjelly.activation_based.base).

Memories (e.g. the membrane potential ``v``) are plain attributes outside ``state_dict``;
``reset()`` restores their registered initial values; ``forward`` dispatches on ``step_mode``.
"""
import copy
import torch
from torch import nn


class StepModule:
    _modes = ("s", "m")

    def supported_step_mode(self):
        return self._modes

    @property
    def step_mode(self):
        return self._step_mode

    @step_mode.setter
    def step_mode(self, value):
        if value not in self.supported_step_mode():
            raise ValueError(f'step_mode can only be {self.supported_step_mode()}, but got "{value}"!')
        self._step_mode = value


class SingleModule(StepModule):
    _modes = ("s",)


class MultiStepModule(StepModule):
    _modes = ("m",)

    @property
    def step_mode(self):
        return "m"

    @step_mode.setter
    def step_mode(self, value):
        if value != "m":
            raise ValueError(f'step_mode can only be ("m",), but got "{value}"!')


class MemoryModule(nn.Module, StepModule):
    def __init__(self):
        super().__init__()
        self._memories, self._memories_rv = {}, {}
        self._backend, self._step_mode = "torch", "s"

    # 'cupy' is accepted so that functional.set_backend(model, 'cupy', ...) of the reference scripts is a
    # no-op: there is a single implementation here, the sm_100a kernels.
    @property
    def supported_backends(self):
        return ("torch", "cupy", "sdf_b200")

    @property
    def backend(self):
        return self._backend

    @backend.setter
    def backend(self, value):
        if value not in self.supported_backends:
            raise NotImplementedError(f"{value} is not a supported backend of {self._get_name()}!")
        self._backend = value

    def single_step_forward(self, x, *a, **k):
        raise NotImplementedError

    def multi_step_forward(self, x_seq, *a, **k):
        return torch.stack([self.single_step_forward(x_seq[t], *a, **k) for t in range(x_seq.shape[0])])

    def forward(self, *a, **k):
        if self.step_mode == "s":
            return self.single_step_forward(*a, **k)
        return self.multi_step_forward(*a, **k)

    def extra_repr(self):
        return f"step_mode={self.step_mode}, backend={self.backend}"

    def register_memory(self, name, value):
        assert not hasattr(self, name), f"{name} has been set as a member variable!"
        self._memories[name] = value
        self._memories_rv[name] = copy.deepcopy(value)

    def reset(self):
        for key in self._memories:
            self._memories[key] = copy.deepcopy(self._memories_rv[key])

    def set_reset_value(self, name, value):
        self._memories_rv[name] = copy.deepcopy(value)

    def __getattr__(self, name):
        mem = self.__dict__.get("_memories")
        if mem is not None and name in mem:
            return mem[name]
        return super().__getattr__(name)

    def __setattr__(self, name, value):
        mem = self.__dict__.get("_memories")
        if mem is not None and name in mem:
            mem[name] = value
        else:
            super().__setattr__(name, value)

    def __delattr__(self, name):
        if name in self._memories:
            del self._memories[name]
            del self._memories_rv[name]
        else:
            super().__delattr__(name)

    def memories(self):
        return iter(self._memories.values())

    def named_memories(self):
        return iter(self._memories.items())

    def detach(self):
        for v in self._memories.values():
            if isinstance(v, torch.Tensor):
                v.detach_()

    def _apply(self, fn, *a, **k):
        for key, v in self._memories.items():
            if isinstance(v, torch.Tensor):
                self._memories[key] = fn(v)
        return super()._apply(fn, *a, **k)
