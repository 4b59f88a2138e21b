"""code/inference_dataset.py:6-25: a folder of images -> transformed tensors."""
import os

from torch.utils.data import Dataset

_EXT = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".tiff")


def make_dataset(root):
    return sorted(os.path.join(d, f) for d, _, fs in os.walk(root) for f in fs if f.lower().endswith(_EXT))


class InferenceDataset(Dataset):
    def __init__(self, root, opts, transform=None, preprocess=None):
        self.paths, self.transform, self.preprocess, self.opts = make_dataset(root), transform, preprocess, opts

    def __len__(self):
        return len(self.paths)

    def __getitem__(self, index):
        from PIL import Image
        p = self.paths[index]
        img = self.preprocess(p) if self.preprocess is not None else Image.open(p).convert("RGB")
        return self.transform(img) if self.transform else img
