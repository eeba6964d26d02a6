"""Encoder networks of the consumer side (idelucs/PytorchUtils.py).  Dense contractions: they
stay PyTorch/cuBLAS by the scope contract; only the layer shapes are taken from the reference
(NetLinear: Linear(4^k,512)-ReLU-Dropout(.5)-Linear(512,64) -> latent;
ReLU-Dropout(.5)-Linear(64,C)-Softmax -> cluster probabilities, PytorchUtils.py:33-56;
myNet, the model_size='small' encoder over the canonical k-mers: Linear(R,400)-ReLU-Dropout(.5)-Linear(400,128)-LeakyReLU,
latent = Linear(128,64), probabilities = Dropout(.5)-Linear(128,C)-Softmax, PytorchUtils.py:6-31)."""
import torch.nn as nn


class NetLinear(nn.Module):
    def __init__(self, n_input, n_output):
        super().__init__()
        self.n_input = n_input
        self.layers = nn.Sequential(nn.Linear(n_input, 512), nn.ReLU(), nn.Dropout(p=0.5), nn.Linear(512, 64))
        self.classifier = nn.Sequential(nn.ReLU(), nn.Dropout(p=0.5), nn.Linear(64, n_output), nn.Softmax(dim=1))

    def forward(self, x):
        latent = self.layers(x.view(-1, self.n_input))
        return self.classifier(latent), latent


class myNet(nn.Module):
    def __init__(self, n_input, n_output):
        super().__init__()
        self.n_input = n_input
        self.layers = nn.Sequential(nn.Linear(n_input, 400), nn.ReLU(), nn.Dropout(p=0.5), nn.Linear(400, 128), nn.LeakyReLU())
        self.instance = nn.Linear(128, 64)
        self.classifier = nn.Sequential(nn.Dropout(p=0.5), nn.Linear(128, n_output), nn.Softmax(dim=1))

    def forward(self, x):
        x = self.layers(x.view(-1, self.n_input))
        return self.classifier(x), self.instance(x)
