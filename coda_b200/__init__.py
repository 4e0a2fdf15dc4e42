"""coda_b200: the CODA active-model-selection acquisition hot path on B200 (sm_100a)."""
from .base import ModelSelector
from .datasets import Dataset, ShardedFileDataset, SyntheticDataset, TensorDataset
from .oracle import Oracle
from .selector import CODA

__all__ = ["CODA", "Dataset", "Oracle", "ModelSelector", "TensorDataset", "SyntheticDataset", "ShardedFileDataset"]
