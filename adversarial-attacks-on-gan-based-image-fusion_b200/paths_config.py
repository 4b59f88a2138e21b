"""code/paths_config.py keys (dataset_paths, model_paths).  Values are empty by default: the reference hard-codes
/home/sh/... paths that do not exist here; fill them to point at real data / checkpoints."""
dataset_paths = {"ffhq": "", "celeba_test": "", "cars_train": "", "cars_test": "", "church_train": "", "church_test": ""}
model_paths = {"ir_se50": "", "stylegan_ffhq": "", "stylegan_cars": "", "stylegan_church": "", "shape_predictor": "", "moco": "",
               "vgg16": ""}
