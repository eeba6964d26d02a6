"""TEST / BENCH INFRASTRUCTURE ONLY — CPU restatement of the reference's training step, used as the reported CPU baseline
of `train.pairs_per_s` (bench.py) and pinned against the live reference in tests/test_oracle_pinning.py.  Never imported by
the product package.

Follows, statement for statement:
  NetLinear            idelucs/PytorchUtils.py:33-56
  weights_init         idelucs/models.py:36-44
  compute_joint        idelucs/LossFunctions.py:49-62
  IID_loss             idelucs/LossFunctions.py:20-46
  info_nce_loss        idelucs/LossFunctions.py:65-98
  AugmentedDataset     idelucs/utils.py:370-389
  create_dataloader    idelucs/utils.py:422-429 (DataLoader(shuffle=True, num_workers=4))
  training epoch       idelucs/models.py:113-143 (zero_grad, two forwards, (1-w) InfoNCE + w IIC, backward, RMSprop step)
with torch on the CPU (the reference's `device` is 'cpu' when no GPU is present, LossFunctions.py:13-17).
"""
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.data import DataLoader, Dataset


class NetLinear(nn.Module):
    def __init__(self, n_input, n_output):
        super().__init__()
        self.n_input = n_input
        self.layers = nn.Sequential(nn.Linear(n_input, 512), nn.ReLU(), nn.Dropout(p=0.5), nn.Linear(512, 64))
        self.classifier = nn.Sequential(nn.ReLU(), nn.Dropout(p=0.5), nn.Linear(64, n_output), nn.Softmax(dim=1))

    def forward(self, x):
        x = x.view(-1, self.n_input)
        latent = self.layers(x)
        return self.classifier(latent), latent


def weights_init(m):
    if isinstance(m, nn.Linear):
        torch.nn.init.kaiming_normal_(m.weight)
        torch.nn.init.zeros_(m.bias)


def compute_joint(x_out, x_tf_out):
    p_i_j = x_out.unsqueeze(2) * x_tf_out.unsqueeze(1)
    p_i_j = p_i_j.sum(dim=0)
    p_i_j = (p_i_j + p_i_j.t()) / 2.
    return p_i_j / p_i_j.sum()


def IID_loss(x_out, x_tf_out, lamb=1.0, EPS=sys.float_info.epsilon):
    _, k = x_out.size()
    p_i_j = compute_joint(x_out, x_tf_out)
    p_i = p_i_j.sum(dim=1).view(k, 1).expand(k, k).clone()
    p_j = p_i_j.sum(dim=0).view(1, k).expand(k, k).clone()
    p_i_j[(p_i_j < EPS).data] = EPS
    p_j[(p_j < EPS).data] = EPS
    p_i[(p_i < EPS).data] = EPS
    loss = - p_i_j * (torch.log(p_i_j) - lamb * torch.log(p_j) - lamb * torch.log(p_i))
    return loss.sum()


def info_nce_loss(z1, z2, temperature):
    N = z1.shape[0]
    features = torch.cat((z1, z2), 0).float()
    labels = torch.cat([torch.arange(N) for _ in range(2)], dim=0)
    labels = (labels.unsqueeze(0) == labels.unsqueeze(1)).float()
    features = F.normalize(features, dim=1)
    similarity_matrix = torch.matmul(features, features.T)
    mask = torch.eye(labels.shape[0], dtype=torch.bool)
    labels = labels[~mask].view(labels.shape[0], -1)
    similarity_matrix = similarity_matrix[~mask].view(similarity_matrix.shape[0], -1)
    positives = similarity_matrix[labels.bool()].view(labels.shape[0], -1)
    negatives = similarity_matrix[~labels.bool()].view(similarity_matrix.shape[0], -1)
    logits = torch.cat([positives, negatives], dim=1)
    labels = torch.zeros(logits.shape[0]).type(torch.LongTensor)
    logits = logits / temperature
    return nn.CrossEntropyLoss()(logits, labels)


class AugmentedDataset(Dataset):
    def __init__(self, data):
        self.data = data

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        return {'true': self.data[idx, 0, :], 'modified': self.data[idx, 1, :]}


class Trainer(object):
    """the parts of idelucs/models.py:46-143 a training epoch touches (model_size 'linear', RMSprop, no scheduler)"""

    def __init__(self, x_train, n_features, n_clusters, batch_sz=512, lamb=2.8, weight=0.25, lr=1e-3, num_workers=4):
        self.n_features = n_features
        self.net = NetLinear(n_features, n_clusters)
        self.net.apply(weights_init)
        self.optimizer = torch.optim.RMSprop(self.net.parameters(), lr=lr, weight_decay=0.01)
        self.l, self.weight = lamb, weight
        self.dataloader = DataLoader(AugmentedDataset(x_train), batch_size=batch_sz, shuffle=True, num_workers=num_workers)

    def contrastive_training_epoch(self):
        self.net.train()
        running_loss = 0.0
        i_batch = 0
        for i_batch, sample_batched in enumerate(self.dataloader):
            sample = sample_batched['true'].view(-1, 1, self.n_features).type(torch.FloatTensor)
            modified_sample = sample_batched['modified'].view(-1, 1, self.n_features).type(torch.FloatTensor)
            self.optimizer.zero_grad()
            z1, h1 = self.net(sample)
            z2, h2 = self.net(modified_sample)
            loss = (1 - self.weight) * info_nce_loss(h1, h2, 0.85) + self.weight * IID_loss(z1, z2, lamb=self.l)
            loss.backward()
            self.optimizer.step()
            running_loss += loss
        running_loss /= i_batch
        return running_loss.item()
