"""``subgraph_counting/lightning_data.py:59-100`` (``LightningDataLoader``): batches for the model entry points.

The reference wraps PyG ``DataLoader``s (forked workers collating ``HeteroData``); here the datasets of
``desco_b200.workload`` batch themselves on the device - consecutive chunks, or a fresh permutation per pass with
``shuffle=True`` (``NeighborhoodBatch.select`` gathers the chunk into its own packed batch) - so ``num_workers`` is accepted
and unused: the CUDA transform cannot run in forked workers.  pytorch_lightning is optional."""
from __future__ import annotations


class LightningDataLoader:
    def __init__(self, train_dataset=None, test_dataset=None, val_dataset=None, batch_size: int = 64,
                 num_workers: int = 0, shuffle: bool = False, generator=None):
        self.train_dataset, self.test_dataset, self.val_dataset = train_dataset, test_dataset, val_dataset
        self.batch_size, self.num_workers, self.shuffle, self.generator = batch_size, num_workers, shuffle, generator

    def _loader(self, ds):
        if ds is None:
            raise ValueError("dataset not set")
        try:
            return ds.loader(self.batch_size, shuffle=self.shuffle, generator=self.generator)
        except TypeError:  # GossipDataset: one batch for the whole dataset, nothing to shuffle
            return ds.loader(self.batch_size)

    def train_dataloader(self):
        return self._loader(self.train_dataset)

    def val_dataloader(self):
        return self._loader(self.val_dataset)

    def test_dataloader(self):
        return self._loader(self.test_dataset)

    predict_dataloader = test_dataloader
