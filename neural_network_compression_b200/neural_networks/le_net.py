"""LeNet300-100 and LeNet5 with the layer shapes of the reference's Keras models
(neural_network_compression/neural_networks/le_net_300_100.py:6-34, le_net_5.py:6-55).  `get_config()` returns
{name: layer} like the reference's, which is what `Trainer.quantize` iterates (trainer.py:50)."""
import torch
from torch import nn


class LeNet300100(nn.Module):
    def __init__(self):
        super().__init__()
        self.dense1 = nn.Linear(784, 300)
        self.dense2 = nn.Linear(300, 100)
        self.out = nn.Linear(100, 10)

    def forward(self, x):
        x = torch.flatten(x, 1)
        return self.out(torch.relu(self.dense2(torch.relu(self.dense1(x)))))

    def get_config(self):
        return {"dense1": self.dense1, "dense2": self.dense2, "out": self.out}

    def layers_to_prune_with_threshold(self):
        # le_net_300_100_trainer.py:22-27
        return {self.dense1: (1, 0.1), self.dense2: (1, 0.1), self.out: (0.5, 0)}


class LeNet5(nn.Module):
    def __init__(self):
        super().__init__()
        # "same" padding and a 2x2 max-pool after each convolution: 28 -> 14 -> 7, 7 * 7 * 50 = 2450 (le_net_5.py:17-34)
        self.conv1 = nn.Conv2d(1, 20, 5, padding=2)
        self.conv2 = nn.Conv2d(20, 50, 5, padding=2)
        self.dense = nn.Linear(2450, 256)
        self.logits = nn.Linear(256, 10)
        self.pool = nn.MaxPool2d(2, 2)

    def forward(self, x):
        x = self.pool(torch.relu(self.conv1(x)))
        x = self.pool(torch.relu(self.conv2(x)))
        return self.logits(torch.relu(self.dense(torch.flatten(x, 1))))

    def get_config(self):
        return {"conv1": self.conv1, "conv2": self.conv2, "dense": self.dense, "logits": self.logits}

    def layers_to_prune_with_threshold(self):
        # no LeNet5 trainer exists in the reference (README.md:140); thresholds of papers/lat/report.tex:252-259
        return {self.conv1: (1, 0.1), self.conv2: (1, 0.1), self.dense: (1, 0.1), self.logits: (1, 0.1)}
