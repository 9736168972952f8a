"""TEST INFRASTRUCTURE ONLY: the reference's train / validation loops restated (sparch/exp.py:341-403, 405-459).

The reference's ``Experiment`` (sparch/exp.py) owns the loop a user of sparch actually runs; it cannot travel to the GPU
box and ``sparch_b200.SNN`` does not run on the CPU, so the drop-in claim of INTEGRATION.md is tested in two halves:
``tests/test_host_contract.py`` runs THESE functions and the reference's own unbound ``Experiment.train_one_epoch`` /
``valid_one_epoch`` on the same CPU model and stub loaders here in the build container (identical losses, learning
rates and parameters afterwards: the restatement is pinned), and ``tests/test_gpu_parity.py`` runs these functions on
the B200 with ``sparch_b200.SNN`` in the place of ``self.net``.  Nothing under ``sparch_b200/`` imports this file.

``self`` is whatever object carries the attributes the reference's methods read: net, train_loader / valid_loader
(yielding ``(x, _, y)``), device, loss_fn, opt, scheduler, use_regularizers, reg_factor, reg_fmin, reg_fmax, save_best,
checkpoint_dir.  Returned values and side effects are the reference's; its ``logging.info`` lines are collected in
``self.log`` (a list) when present.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _log(self, msg):
    if hasattr(self, "log"):
        self.log.append(msg)


def train_one_epoch(self, e):
    """sparch/exp.py:341-403."""
    self.net.train()                                            # exp.py:347
    losses, accs = [], []
    epoch_spike_rate = 0
    for step, (x, _, y) in enumerate(self.train_loader):        # exp.py:352
        x = x.to(self.device)                                   # exp.py:355-356
        y = y.to(self.device)
        output, firing_rates = self.net(x)                      # exp.py:359
        loss_val = self.loss_fn(output, y)                      # exp.py:362
        losses.append(loss_val.item())                          # exp.py:363 (host sync)
        if self.net.is_snn:                                     # exp.py:366-372
            epoch_spike_rate += torch.mean(firing_rates)
            if self.use_regularizers:
                reg_quiet = F.relu(self.reg_fmin - firing_rates).sum()
                reg_burst = F.relu(firing_rates - self.reg_fmax).sum()
                loss_val += self.reg_factor * (reg_quiet + reg_burst)
        self.opt.zero_grad()                                    # exp.py:375-377
        loss_val.backward()
        self.opt.step()
        pred = torch.argmax(output, dim=1)                      # exp.py:380-382
        acc = np.mean((y == pred).detach().cpu().numpy())
        accs.append(acc)
    current_lr = self.opt.param_groups[-1]["lr"]                # exp.py:385
    _log(self, f"Epoch {e}: lr={current_lr}")
    train_loss = np.mean(losses)                                # exp.py:389
    _log(self, f"Epoch {e}: train loss={train_loss}")
    train_acc = np.mean(accs)
    _log(self, f"Epoch {e}: train acc={train_acc}")
    if self.net.is_snn:                                         # exp.py:397-399 (divides by the LAST step index)
        epoch_spike_rate /= step
        _log(self, f"Epoch {e}: train mean act rate={epoch_spike_rate}")
    return losses, accs


def valid_one_epoch(self, e, best_epoch, best_acc):
    """sparch/exp.py:405-459."""
    with torch.no_grad():
        self.net.eval()                                         # exp.py:412
        losses, accs = [], []
        epoch_spike_rate = 0
        for step, (x, _, y) in enumerate(self.valid_loader):
            x = x.to(self.device)
            y = y.to(self.device)
            output, firing_rates = self.net(x)
            loss_val = self.loss_fn(output, y)
            losses.append(loss_val.item())
            pred = torch.argmax(output, dim=1)
            acc = np.mean((y == pred).detach().cpu().numpy())
            accs.append(acc)
            if self.net.is_snn:
                epoch_spike_rate += torch.mean(firing_rates)
        valid_loss = np.mean(losses)
        _log(self, f"Epoch {e}: valid loss={valid_loss}")
        valid_acc = np.mean(accs)
        _log(self, f"Epoch {e}: valid acc={valid_acc}")
        if self.net.is_snn:
            epoch_spike_rate /= step
        self.scheduler.step(valid_acc)                          # exp.py:447
        if valid_acc > best_acc:                                # exp.py:450-457
            best_acc = valid_acc
            best_epoch = e
            if self.save_best:
                torch.save(self.net, f"{self.checkpoint_dir}/best_model.pth")
        return best_epoch, best_acc


def stub_experiment(net, device, lr=1e-2, batches=4, batch=8, T=20, F_in=40, classes=10, seed=0, **over):
    """An object with the attributes the loops read, built the way Experiment.__init__ builds them (exp.py:89-100):
    torch.optim.Adam, ReduceLROnPlateau(mode="max", min_lr=1e-6), nn.CrossEntropyLoss, and two lists of
    ``(x, lengths, y)`` CPU batches standing in for the DataLoaders."""
    import types
    from torch.optim.lr_scheduler import ReduceLROnPlateau
    g = torch.Generator().manual_seed(seed)
    proto = torch.randn(classes, T, F_in, generator=g)          # class prototypes: the task is learnable
    def loader(n):
        out = []
        for _ in range(n):
            y = torch.randint(0, classes, (batch,), generator=g)
            x = proto[y] + 0.5 * torch.randn(batch, T, F_in, generator=g)
            out.append((x, torch.full((batch,), T), y))
        return out
    self = types.SimpleNamespace(
        net=net, device=device, train_loader=loader(batches), valid_loader=loader(2),
        loss_fn=torch.nn.CrossEntropyLoss(), use_regularizers=False, reg_factor=0.5, reg_fmin=0.01, reg_fmax=0.5,
        save_best=False, checkpoint_dir=None, log=[])
    self.opt = torch.optim.Adam(net.parameters(), lr)
    self.scheduler = ReduceLROnPlateau(optimizer=self.opt, mode="max", factor=0.7, patience=1, min_lr=1e-6)
    for k, v in over.items():
        setattr(self, k, v)
    return self
