"""Seed generators of the sampling CLI (reference sampler/random_util.py), restated.

``get_generator(kind, num_samples, seed)`` -> object with ``randn(*size, dtype, device)``,
``randint(low, high, size, dtype, device)``, ``randn_like(t)``:

* ``dummy``         plain torch RNG (random_util.py:25-33)
* ``determ``        ONE generator draws the whole ``[num_samples, ...]`` population and the caller receives rows
                    ``done + rank, done + rank + world, ...`` (clamped) - batch-size / world-size independent noise
                    (random_util.py:36-96)
* ``determ-indiv``  one generator per sample index, seeded ``i + num_samples * seed`` (random_util.py:99-173)
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class DummyGenerator:
    def randn(self, *a, **k):
        return torch.randn(*a, **k)

    def randint(self, *a, **k):
        return torch.randint(*a, **k)

    def randn_like(self, *a, **k):
        return torch.randn_like(*a, **k)


class _IndexedBase:
    def __init__(self, num_samples, seed):
        self.rank, self.world_size = _rank_world()
        self.num_samples, self.done_samples, self.seed = num_samples, 0, seed

    def _indices(self, n):
        idx = torch.arange(self.done_samples + self.rank, self.done_samples + self.world_size * int(n), self.world_size)
        idx = idx.clamp(0, self.num_samples - 1)
        assert len(idx) == n
        return idx

    def randn_like(self, t):
        return self.randn(*t.size(), dtype=t.dtype, device=t.device)

    def get_seed(self):
        return self.seed


class DeterministicGenerator(_IndexedBase):
    def __init__(self, num_samples, seed=0):
        super().__init__(num_samples, seed)
        self.rng_cpu = torch.Generator()
        self.rng_cuda = torch.Generator("cuda") if torch.cuda.is_available() else None
        self.set_seed(seed)

    def _gen(self, device):
        return self.rng_cpu if torch.device(device).type == "cpu" else self.rng_cuda

    def randn(self, *size, dtype=torch.float, device="cpu"):
        pop = torch.randn(self.num_samples, *size[1:], generator=self._gen(device), dtype=dtype, device=device)
        return pop[self._indices(size[0]).to(pop.device)]

    def randint(self, low, high, size, dtype=torch.long, device="cpu"):
        pop = torch.randint(low, high, generator=self._gen(device), size=(self.num_samples, *size[1:]), dtype=dtype,
                            device=device)
        return pop[self._indices(size[0]).to(pop.device)]

    def set_done_samples(self, done):
        self.done_samples = done
        self.set_seed(self.seed)

    def set_seed(self, seed):
        self.rng_cpu.manual_seed(seed)
        if self.rng_cuda is not None:
            self.rng_cuda.manual_seed(seed)


class DeterministicIndividualGenerator(_IndexedBase):
    def __init__(self, num_samples, seed=0):
        super().__init__(num_samples, seed)
        self.rng_cpu = [torch.Generator() for _ in range(num_samples)]
        self.rng_cuda = [torch.Generator("cuda") for _ in range(num_samples)] if torch.cuda.is_available() else None
        self.set_seed(seed)

    def _gen(self, device):
        return self.rng_cpu if torch.device(device).type == "cpu" else self.rng_cuda

    def randn(self, *size, dtype=torch.float, device="cpu"):
        gens = self._gen(device)
        return torch.cat([torch.randn(1, *size[1:], generator=gens[i], dtype=dtype, device=device)
                          for i in self._indices(size[0]).tolist()], dim=0)

    def randint(self, low, high, size, dtype=torch.long, device="cpu"):
        gens = self._gen(device)
        return torch.cat([torch.randint(low, high, generator=gens[i], size=(1, *size[1:]), dtype=dtype, device=device)
                          for i in self._indices(size[0]).tolist()], dim=0)

    def set_done_samples(self, done):
        self.done_samples = done

    def set_seed(self, seed):
        for i, g in enumerate(self.rng_cpu):
            g.manual_seed(i + self.num_samples * seed)
        if self.rng_cuda is not None:
            for i, g in enumerate(self.rng_cuda):
                g.manual_seed(i + self.num_samples * seed)


def get_generator(generator, num_samples=0, seed=0):
    if generator == "dummy":
        return DummyGenerator()
    if generator == "determ":
        return DeterministicGenerator(num_samples, seed)
    if generator == "determ-indiv":
        return DeterministicIndividualGenerator(num_samples, seed)
    raise NotImplementedError(generator)
