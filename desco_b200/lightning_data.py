"""``subgraph_counting/lightning_data.py:59-100`` (``LightningDataLoader``): batches for the model entry points.
pytorch_lightning / PyG loaders are optional here - the datasets of ``desco_b200.workload`` batch themselves."""
from __future__ import annotations


class LightningDataLoader:
    def __init__(self, train_dataset=None, test_dataset=None, val_dataset=None, batch_size: int = 64,
                 num_workers: int = 0, shuffle: bool = False):
        if shuffle:
            raise NotImplementedError("shuffle=True is a training option; the inference hot path iterates in order")
        self.train_dataset, self.test_dataset, self.val_dataset = train_dataset, test_dataset, val_dataset
        self.batch_size, self.num_workers, self.shuffle = batch_size, num_workers, shuffle  # workers: GPU path, unused

    def _loader(self, ds):
        if ds is None:
            raise ValueError("dataset not set")
        return ds.loader(self.batch_size)

    def train_dataloader(self):
        return self._loader(self.train_dataset)

    def val_dataloader(self):
        return self._loader(self.val_dataset)

    def test_dataloader(self):
        return self._loader(self.test_dataset)

    predict_dataloader = test_dataloader
