"""Host-side scalar schedules of src/utils/schedulers.py (WarmupCosineSchedule :41-68,
CosineWDSchedule :71-93) and the momentum generator of app/vjepa/train.py:286-289.  They only produce
floats; the fused optimizer kernels take them as arguments."""
import math


class WarmupCosineSchedule(object):
    def __init__(self, warmup_steps, start_lr, ref_lr, T_max, final_lr=0.0):
        self.start_lr, self.ref_lr, self.final_lr = start_lr, ref_lr, final_lr
        self.warmup_steps = warmup_steps
        self.T_max = T_max - warmup_steps
        self._step = 0.0

    def step(self):
        self._step += 1
        if self._step < self.warmup_steps:
            progress = float(self._step) / float(max(1, self.warmup_steps))
            return self.start_lr + progress * (self.ref_lr - self.start_lr)
        progress = float(self._step - self.warmup_steps) / float(max(1, self.T_max))
        return max(self.final_lr,
                   self.final_lr + (self.ref_lr - self.final_lr) * 0.5 * (1.0 + math.cos(math.pi * progress)))


class CosineWDSchedule(object):
    def __init__(self, ref_wd, T_max, final_wd=0.0):
        self.ref_wd, self.final_wd, self.T_max = ref_wd, final_wd, T_max
        self._step = 0.0

    def step(self):
        self._step += 1
        progress = self._step / self.T_max
        new_wd = self.final_wd + (self.ref_wd - self.final_wd) * 0.5 * (1.0 + math.cos(math.pi * progress))
        return max(self.final_wd, new_wd) if self.final_wd <= self.ref_wd else min(self.final_wd, new_wd)


def momentum_schedule(ema, ipe, num_epochs, ipe_scale):
    """train.py:286-289."""
    return (ema[0] + i * (ema[1] - ema[0]) / (ipe * num_epochs * ipe_scale)
            for i in range(int(ipe * num_epochs * ipe_scale) + 1))
