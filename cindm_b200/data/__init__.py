from .nbody_dataset import NBodyDataset, first_batch_1d, get_item_1d  # noqa: F401
