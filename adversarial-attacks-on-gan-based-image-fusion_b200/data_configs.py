"""code/data_configs.py: DATASETS[name] -> {transforms, *_root}."""
from . import paths_config as _p
from . import transforms_config

DATASETS = {
    "ffhq_encode": {"transforms": transforms_config.EncodeTransforms, "train_source_root": _p.dataset_paths["ffhq"],
                    "train_target_root": _p.dataset_paths["ffhq"], "test_source_root": _p.dataset_paths["celeba_test"],
                    "test_target_root": _p.dataset_paths["celeba_test"]},
    "cars_encode": {"transforms": transforms_config.CarsEncodeTransforms, "train_source_root": _p.dataset_paths["cars_train"],
                    "train_target_root": _p.dataset_paths["cars_train"], "test_source_root": _p.dataset_paths["cars_test"],
                    "test_target_root": _p.dataset_paths["cars_test"]},
    "church_encode": {"transforms": transforms_config.EncodeTransforms, "train_source_root": _p.dataset_paths["church_train"],
                      "train_target_root": _p.dataset_paths["church_train"], "test_source_root": _p.dataset_paths["church_test"],
                      "test_target_root": _p.dataset_paths["church_test"]},
}
