import random

from .base_tuner import BaseTuner


class RandomTuner(BaseTuner):

    def next_batch(self, sample_size=1):
        out = []
        for _ in range(min(sample_size, len(self.all_exps))):
            out.append(self.all_exps.pop(random.randrange(len(self.all_exps))))
        return out


class GridSearchTuner(BaseTuner):

    def next_batch(self, sample_size=1):
        out = self.all_exps[:sample_size]
        self.all_exps = self.all_exps[sample_size:]
        return out
