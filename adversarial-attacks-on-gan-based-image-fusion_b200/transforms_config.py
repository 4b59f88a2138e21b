"""code/transforms_config.py: Resize -> ToTensor -> Normalize([0.5]*3,[0.5]*3) (images in [-1,1], :28-31,60-63)."""
from abc import abstractmethod


class TransformsConfig(object):
    def __init__(self, opts):
        self.opts = opts

    @abstractmethod
    def get_transforms(self):
        pass


def _t(size):
    import torchvision.transforms as transforms
    return transforms.Compose([transforms.Resize(size), transforms.ToTensor(), transforms.Normalize([0.5] * 3, [0.5] * 3)])


class EncodeTransforms(TransformsConfig):
    def get_transforms(self):
        s = getattr(self.opts, "stylegan_size", 1024) if self.opts is not None else 1024
        return {"transform_gt_train": _t((s, s)), "transform_source": None, "transform_test": _t((s, s)), "transform_inference": _t((s, s))}


class CarsEncodeTransforms(TransformsConfig):
    def get_transforms(self):
        return {"transform_gt_train": _t((384, 512)), "transform_source": None, "transform_test": _t((384, 512)),
                "transform_inference": _t((384, 512))}
