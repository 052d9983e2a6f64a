"""Index samplers (drop-in for reference dataloader/sampler.py:6-79)."""
import torch
from torch.utils.data.sampler import Sampler


class EpochConcateSampler(Sampler):
    """Concatenation of `epoch` independent shuffles of the dataset indices."""

    def __init__(self, data_source, epoch):
        self.data_length, self.epoch = len(data_source), epoch

    def __iter__(self):
        idx = []
        for _ in range(self.epoch):
            idx += torch.randperm(self.data_length).tolist()
        return iter(idx)

    def __len__(self):
        return self.data_length * self.epoch


class EpochConcateDistributedSampler(Sampler):
    """Rank-sharded variant: every rank draws the same shuffles (shared seed) and keeps indices rank::world."""

    def __init__(self, data_source, epoch, rank, world_size, seed=0):
        self.n, self.epoch, self.rank, self.world, self.seed = len(data_source), epoch, rank, world_size, seed

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        idx = []
        for _ in range(self.epoch):
            idx += torch.randperm(self.n, generator=g).tolist()
        usable = (len(idx) // self.world) * self.world
        return iter(idx[self.rank:usable:self.world])

    def __len__(self):
        return (self.n * self.epoch) // self.world
