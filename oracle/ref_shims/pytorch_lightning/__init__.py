import torch.nn as nn
LightningModule = nn.Module
