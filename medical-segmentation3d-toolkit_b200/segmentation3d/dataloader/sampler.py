"""Index samplers (drop-in for reference dataloader/sampler.py:6-79): every epoch's shuffle concatenated into one index
stream, drawn from the same generators as the reference (python's `random` for the single-process samplers,
DistributedSampler's seeded torch generator for the rank-sharded one)."""
import random

from torch.utils.data.distributed import DistributedSampler
from torch.utils.data.sampler import Sampler


class EpochConcateSampler(Sampler):
    """`epoch` independent `random.shuffle`s of range(len(data_source)) (:6-26)."""

    def __init__(self, data_source, epoch):
        self.data_length = len(data_source)
        self.epoch = epoch

    def __iter__(self):
        index_all = []
        for _ in range(self.epoch):
            index = list(range(self.data_length))
            random.shuffle(index)
            index_all += index
        return iter(index_all)

    def __len__(self):
        return self.data_length * self.epoch


class EpochConcateSamplerResume(Sampler):
    """same, with epoch i shuffled under random.seed(i) so a resumed run continues the stream (:29-53)."""

    def __init__(self, data_source, epoch, resume_epoch):
        self.data_length = len(data_source)
        self.epoch = epoch
        self.resume_epoch = resume_epoch

    def __iter__(self):
        index_all = []
        for i in range(self.resume_epoch, self.resume_epoch + self.epoch):
            index = list(range(self.data_length))
            random.seed(i)
            random.shuffle(index)
            index_all += index
        return iter(index_all)

    def __len__(self):
        return self.data_length * self.epoch


class EpochConcateDistributedSampler(DistributedSampler):
    """rank's share of every epoch's DistributedSampler shuffle, concatenated (:56-79).  `rank` / `world_size` default to
    the initialised process group like the reference; passing them explicitly needs no process group."""

    def __init__(self, data_source, epoch, resume_epoch=0, rank=None, world_size=None, seed=0):
        super(EpochConcateDistributedSampler, self).__init__(data_source, num_replicas=world_size, rank=rank, seed=seed)
        self.data_length = len(data_source)
        # the reference keeps the epoch COUNT in `self.epoch`, which DistributedSampler.set_epoch then overwrites with the
        # current epoch (so its __len__ changes while iterating); the count lives in its own attribute here
        self.num_epochs = epoch
        self.resume_epoch = resume_epoch

    def __iter__(self):
        index_all = []
        for i in range(self.resume_epoch, self.resume_epoch + self.num_epochs):
            super(EpochConcateDistributedSampler, self).set_epoch(i)
            index_all += list(super(EpochConcateDistributedSampler, self).__iter__())
        return iter(index_all)

    def __len__(self):
        return super(EpochConcateDistributedSampler, self).__len__() * self.num_epochs
