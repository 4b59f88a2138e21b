"""Run artefacts in the reference's own formats, so that downstream scripts written against the reference's run
folders keep working.  Host-side bookkeeping only; nothing here launches a kernel.

  new_run_folder(dir)                         code/attack/attack_main2.py:407-410
  write_parameters(dir, adversarial, args, size, max_iter)   `parameters.txt`, the `key value` lines of :975-988 (appended)
  RunRecorder                                 the six lists of :990-995 and their `torch.save(torch.cat(...), '<name>.npz')` dumps at
                                              :1098-1111 -- torch pickles despite the .npz suffix, benign/ and adversarial/ sub-folders
  result_columns(n_inputs)                    the xlsx header of code/attack/interpolation.py:1256-1258
  result_row(noise, cri_spati, ...)           one row in that column order, from the dicts cal_result returns (:1076-1091)
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

DATASET_N = {"ffhq": 5, "car": 4, "church": 3}          # code/attack/attack_main2.py:909


def new_run_folder(file_dir: str) -> str:
    if not os.path.exists(file_dir):
        os.mkdir(file_dir)
    return file_dir


def write_parameters(attack_savedir: str, adversarial: str, args, decoder_size: int, max_iter: int) -> str:
    """appends the reference's twelve `key value` lines; returns the file path (the `param_file` global of :975)"""
    param_file = os.path.join(attack_savedir, "parameters.txt")
    g = lambda k: getattr(args, k, None)
    with open(param_file, "a") as f:
        f.write("adversarial attack {}\n".format(adversarial))
        f.write("dataset {}\n".format(g("dataset_name")))
        f.write("dataset size {}\n".format(decoder_size))
        f.write("epochs {}\n".format(g("epochs")))
        f.write("max_count {}\n".format(g("max_count")))
        f.write("patch_size {}\n".format(g("patch_size")))
        f.write("train_size {}\n".format(g("train_size")))
        f.write("patch_type {}\n".format(g("patch_type")))
        f.write("white-box max_iter {}\n".format(max_iter))
        f.write("white-box lr {}\n".format(g("lr")))
        f.write("use_generate_img {}\n".format(g("use_generate_img")))
    return param_file


class RunRecorder:
    """Collects what the reference's attack driver accumulates per batch and writes it the way :1098-1111 does."""

    NAMES = (("adversarial", "all_adv_inputs"), ("benign", "all_inputs"), ("adversarial", "all_adv_rec_loss"), ("benign", "all_rec_loss"),
             ("adversarial", "all_adv_inner_feature"), ("benign", "all_inner_feature"))

    def __init__(self, attack_savedir: str):
        self.benign_savedir = new_run_folder(os.path.join(attack_savedir, "benign"))            # :971-972
        self.adv_savedir = new_run_folder(os.path.join(attack_savedir, "adversarial"))
        self.lists: Dict[str, List[torch.Tensor]] = {name: [] for _, name in self.NAMES}

    def add(self, **tensors: torch.Tensor):
        """add(all_inputs=x, all_adv_inputs=x_adv, all_rec_loss=..., ...): each tensor is detached and moved to the host"""
        for k, v in tensors.items():
            if k not in self.lists:
                raise KeyError(f"{k!r} is not one of {[n for _, n in self.NAMES]}")
            self.lists[k].append(v.detach().cpu())

    def save(self) -> List[str]:
        """torch.save(torch.cat(list, 0), <folder>/<name>.npz) for every non-empty list; returns the paths written"""
        out = []
        for folder, name in self.NAMES:
            if self.lists[name]:
                path = os.path.join(self.adv_savedir if folder == "adversarial" else self.benign_savedir, name + ".npz")
                torch.save(torch.cat(self.lists[name], dim=0), path)
                out.append(path)
        return out


def result_columns(n_inputs: int) -> List[str]:
    n = n_inputs
    return (["noise"] * n + ["cri_spati"] * (n + 1) + ["cri_arith"] * (n + 1) + ["vg_spati"] * (n + 1) + ["vg_arith"] * (n + 1) +
            ["ssmi_spati"] * (n + 1) + ["ssmi_arith"] * (n + 1))


def result_row(noise: Sequence[float], cri_spati: Dict[int, float], cri_arith: Dict[int, float], vg_spati: Dict[int, float],
               vg_arith: Dict[int, float], ssmi_spati: Dict[int, float], ssmi_arith: Dict[int, float]) -> List[float]:
    """One results row in the column order above.  The six dicts are cal_result's returns ({i: value} over the n_inputs + 1
    adversarial fusions, interpolation.py:1076-1091) for the spatial and the arithmetic fusion."""
    n = len(noise)
    row = list(noise)
    for d in (cri_spati, cri_arith, vg_spati, vg_arith, ssmi_spati, ssmi_arith):
        if len(d) != n + 1:
            raise ValueError(f"expected {n + 1} entries per metric (one per adversarial fusion), got {len(d)}")
        row += [float(d[i]) for i in range(n + 1)]
    return row
