"""Training loop of the gain-regressor CNNs around the fused hot path (config 4 of BASELINE.json).

Drop-in for the reference's ``ModelTrainer`` (model_trainer.py:5-68: same constructor, ``_train_epoch``,
``_validate_epoch`` and ``fit``; MSE on the masked dB map, one optimiser step per batch) plus what the
reference does not have:

* ``FeatureBatcher``: the reference featurises inside ``Dataset.__getitem__`` -- five ``compute_features`` calls
  per sample in DataLoader worker processes (data/dataset.py:185-210).  Here a whole batch of raw chunks
  ``[B, 5, C, n]`` (bass, drums, vocals, other, mix: data/dataset.py:42,192-207) resident on the device goes
  through ONE K1 launch that writes the model input ``[B, 4, 1025, T]`` and the target ``[B, 1025, T]`` in place.
* data parallelism: ``ModelTrainer(..., ddp=True)`` wraps the model in ``DistributedDataParallel`` (NCCL on
  GPUs, gloo in the CPU tests); every rank trains on its own shard of the batch and the gradients are
  averaged by all-reduce -- the only collective of the training configuration.

The CNN body stays PyTorch/cuDNN (north_star); its tail ``masked = sum_i m_i x_i`` and the gradient with
respect to ``m`` are kernel K3 (mixing.feature_mix).
"""
import os

import torch

from . import features


class FeatureBatcher:
    """raw chunks -> (train_features, gt_features) on the device, one kernel launch per batch."""

    def __init__(self, window_size: int = 2048, hop_length: int = 1024, compute_dtype=torch.float64):
        # float64 arithmetic reproduces the reference's training path (soundfile yields float64, so its features
        # are float64 and are cast with .float() in front of the model: training_ignite.ipynb:291); the result is
        # rounded to float32 once, exactly like that cast
        self.n_fft, self.hop, self.compute_dtype = window_size, hop_length, compute_dtype

    def __call__(self, raw: torch.Tensor, row_gain: torch.Tensor = None):
        """raw [B, 5, C, n] or [B, 5, n] CUDA (int16 / float32 / float64).  ``row_gain`` [B, 5] is the audio
        augmentation factor of data/dataset.py:164-168 (applied after the mono down-mix, before the transform).
        Returns (train [B, 4, F, T] float32, gt [B, F, T] float32), both views of one [B, 5, F, T] buffer."""
        if raw.dim() not in (3, 4) or raw.shape[1] != 5:
            raise ValueError('raw must be [B, 5, C, n] or [B, 5, n] (bass, drums, vocals, other, mix)')
        B = raw.shape[0]
        kw = {'channel_axis': -2} if raw.dim() == 4 else {}
        rows = raw.reshape((B * 5,) + tuple(raw.shape[2:]))
        gain = row_gain.reshape(B * 5) if row_gain is not None else None
        out = features.stft_db(rows, self.n_fft, self.hop, compute_dtype=self.compute_dtype,
                               out_dtype=torch.float32, row_gain=gain, **kw)
        out = out.reshape((B, 5) + tuple(out.shape[1:]))
        return out[:, :4], out[:, 4]


def wrap_ddp(model: torch.nn.Module, device) -> torch.nn.Module:
    """DistributedDataParallel around ``model`` when a process group with more than one rank exists."""
    dist = torch.distributed
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return model
    device = torch.device(device)
    if device.type == 'cuda':
        return torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], output_device=device.index)
    return torch.nn.parallel.DistributedDataParallel(model)


class ModelTrainer:
    def __init__(self, model, criterion, optimizer, device, model_name='scalar2d', ddp=False, log_every=10):
        self.weights_dir = './weights'
        self.model_name = model_name
        self.model = wrap_ddp(model, device) if ddp else model
        self.optimizer = optimizer
        self.criterion = criterion
        self.device = device
        self.log_every = log_every

    def train_step(self, train_features: torch.Tensor, gt_features: torch.Tensor) -> torch.Tensor:
        """One optimiser step (model_trainer.py:29-38); returns the loss as a device scalar (no host sync)."""
        self.optimizer.zero_grad()
        masked, _ = self.model(train_features.to(self.device))
        loss = self.criterion(masked, gt_features.to(self.device))
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _run_epoch(self, loader, train: bool) -> float:
        total = torch.zeros((), dtype=torch.float64, device=self.device)
        n = 0
        for i, (train_features, gt_features) in enumerate(loader):
            if train:
                loss = self.train_step(train_features, gt_features)
                if self.log_every and i % self.log_every == self.log_every - 1:
                    print('[%d/%4d] loss: %.3f' % (i + 1, len(loader), loss.item()))
            else:
                masked, _ = self.model(train_features.to(self.device))
                loss = self.criterion(masked, gt_features.to(self.device)).detach()
            total += loss.double()
            n += 1
        return float(total.item()) / max(n, 1)

    def _validate_epoch(self, val_loader):
        with torch.no_grad():
            return self._run_epoch(val_loader, train=False)

    def _train_epoch(self, train_loader):
        return self._run_epoch(train_loader, train=True)

    def fit(self, train_loader, val_loader, start_epoch, num_epochs):
        train_loss, val_loss = [], []
        for epoch in range(start_epoch, start_epoch + num_epochs):
            print('Epoch {}/{}'.format(epoch, num_epochs - 1))
            train_loss.append(self._train_epoch(train_loader))
            print('Epoch {} train loss: {:.4f}'.format(epoch, train_loss[-1]))
            val_loss.append(self._validate_epoch(val_loader))
            print('Epoch {} val loss: {:.4f}'.format(epoch, val_loss[-1]))
            print('-' * 50)
            rank0 = not (torch.distributed.is_available() and torch.distributed.is_initialized()) or \
                torch.distributed.get_rank() == 0
            if rank0:
                os.makedirs(self.weights_dir, exist_ok=True)
                inner = self.model.module if hasattr(self.model, 'module') else self.model
                torch.save(inner.state_dict(), os.path.join(
                    self.weights_dir, 'mixmodel_{}_1s_{:04d}_{:.4f}.pt'.format(self.model_name, epoch, train_loss[-1])))
        return train_loss, val_loss
