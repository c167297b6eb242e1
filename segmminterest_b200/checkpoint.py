"""`CheckPointer` of kn_util/nn_utils/checkpoint.py:11-77 as the driver uses it (main_for_seq_leave_earlystop_SegMM.py:217,
333,366): `CheckPointer("main_metric", dir, mode="max", cur_time=...)`, `save_checkpoint(model=, optimizer=, num_epochs=,
metric_vals=)` -> bool (a new best), `load_checkpoint(model, optimizer, mode='best')` -> the saved dict.

Same file names (`ckpt-latest.pth`, `ckpt-best-ep{E}-{metric}.pth`) and the same saved keys, so checkpoints written by
either implementation load in the other (the state_dict schema is the reference's: tests/test_host_logic.py).  Differences:
the shipped class does not accept the `cur_time=` keyword the driver passes (TypeError there) -- accepted and recorded here;
`work_dir` is created when missing; the best checkpoint is found by globbing its own pattern (the shipped code joins
`work_dir` twice, :58, and finds nothing for a relative directory); old best files are removed with os.remove instead of
`rm -rf` through a shell."""
from __future__ import annotations

import glob
import os
import os.path as osp

import numpy as np
import torch


class CheckPointer:
    def __init__(self, monitor, work_dir, mode="min", cur_time=None) -> None:
        self.monitor = monitor
        self.best_metric = None
        self.work_dir = work_dir
        self.mode = mode
        self.cur_time = cur_time
        self.ckpt_latest = osp.join(self.work_dir, "ckpt-latest.pth")
        self.ckpt_best = osp.join(self.work_dir, "ckpt-best-ep{}-{}.pth")

    def better(self, new, orig):
        if orig is None:
            return True
        if self.mode == "min":
            return new < orig
        elif self.mode == "max":
            return new > orig
        raise NotImplementedError()

    def save_checkpoint(self, model, optimizer, num_epochs, metric_vals=None, loss_scaler=None, lr_scheduler=None):
        """latest checkpoint always; the best one when metric_vals[monitor] improves (returns True then)."""
        os.makedirs(self.work_dir, exist_ok=True)
        save_dict = dict(model=model.state_dict(), optimizer=optimizer.state_dict(), num_epochs=num_epochs, metrics=metric_vals)
        if self.cur_time is not None:
            save_dict["cur_time"] = self.cur_time
        if lr_scheduler:
            save_dict["lr_scheduler"] = lr_scheduler.state_dict()
        if loss_scaler:
            save_dict["loss_scaler"] = loss_scaler.state_dict()
        torch.save(save_dict, self.ckpt_latest)
        if metric_vals:
            if self.better(metric_vals[self.monitor], self.best_metric):
                self.best_metric = metric_vals[self.monitor]
                for old in glob.glob(self.ckpt_best.format("*", "*")):
                    os.remove(old)
                torch.save(save_dict, self.ckpt_best.format(num_epochs, np.round(self.best_metric, decimals=6)))
                return True
        return False

    def load_checkpoint(self, model, optimizer, lr_scheduler=None, loss_scaler=None, mode="latest"):
        if mode == "latest":
            fn = self.ckpt_latest
        elif mode == "best":
            fn = glob.glob(self.ckpt_best.format("*", "*"))[0]
        else:
            raise NotImplementedError()
        load_dict = torch.load(fn, weights_only=False)
        model.load_state_dict(load_dict["model"])
        optimizer.load_state_dict(load_dict["optimizer"])
        if lr_scheduler:
            if "lr_scheduler" not in load_dict:
                raise Exception("lr_scheduler not found")
            lr_scheduler.load_state_dict(load_dict["lr_scheduler"])
        if loss_scaler:
            if "loss_scaler" not in load_dict:
                raise Exception("loss_scaler not found")
            loss_scaler.load_state_dict(load_dict["loss_scaler"])
        return load_dict
