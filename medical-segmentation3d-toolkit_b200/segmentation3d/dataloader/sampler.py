"""Index samplers behind the training DataLoader (interface of reference dataloader/sampler.py:6-79).

A sampler here is a recipe "permutation of epoch e" plus the list of epochs to play; the index stream is the
concatenation of those permutations.  The permutations come from the generators the reference uses (python's `random`
for the single-process samplers, DistributedSampler's seeded torch generator for the rank-sharded one), so a seeded
run visits the cases in the reference's order (tests/golden/samplers.json)."""
import itertools
import random

from torch.utils.data.distributed import DistributedSampler
from torch.utils.data.sampler import Sampler


def _python_shuffle(n, seed=None):
    order = list(range(n))
    if seed is not None:
        random.seed(seed)
    random.shuffle(order)
    return order


class _EpochStream(Sampler):
    """concatenation of one permutation per epoch in `self.epochs()`"""

    def __init__(self, data_source, epoch):
        self.data_length, self.epoch = len(data_source), epoch

    def epochs(self):
        return range(self.epoch)

    def permutation(self, e):
        raise NotImplementedError

    def __iter__(self):
        return itertools.chain.from_iterable([self.permutation(e) for e in self.epochs()])

    def __len__(self):
        return self.epoch * self.data_length


class EpochConcateSampler(_EpochStream):
    """every epoch an unseeded `random.shuffle` of all indices (:6-26)"""

    def permutation(self, e):
        return _python_shuffle(self.data_length)


class EpochConcateSamplerResume(_EpochStream):
    """epoch e shuffled under `random.seed(e)`, starting at `resume_epoch`, so a resumed run continues the stream (:29-53)"""

    def __init__(self, data_source, epoch, resume_epoch):
        super().__init__(data_source, epoch)
        self.resume_epoch = resume_epoch

    def epochs(self):
        return range(self.resume_epoch, self.resume_epoch + self.epoch)

    def permutation(self, e):
        return _python_shuffle(self.data_length, seed=e)


class EpochConcateDistributedSampler(DistributedSampler):
    """This rank's share of DistributedSampler's shuffle of every epoch, concatenated (:56-79).  `rank` / `world_size`
    default to the initialised process group; passing both needs no process group.  The number of epochs is kept in
    `num_epochs` (DistributedSampler.set_epoch owns `self.epoch`)."""

    def __init__(self, data_source, epoch, resume_epoch=0, rank=None, world_size=None, seed=0):
        DistributedSampler.__init__(self, data_source, num_replicas=world_size, rank=rank, seed=seed)
        self.data_length, self.num_epochs, self.resume_epoch = len(data_source), epoch, resume_epoch

    def _share(self, e):
        self.set_epoch(e)
        return list(DistributedSampler.__iter__(self))

    def __iter__(self):
        first = self.resume_epoch
        return itertools.chain.from_iterable([self._share(e) for e in range(first, first + self.num_epochs)])

    def __len__(self):
        return self.num_epochs * DistributedSampler.__len__(self)
