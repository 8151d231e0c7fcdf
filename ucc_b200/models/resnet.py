"""ResNet-50 (He et al. 2015) in plain torch.nn: the model of BASELINE.json's DDP configuration."""
import torch
from torch import nn


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, width, stride=1, down=None):
        super().__init__()
        self.c1 = nn.Conv2d(cin, width, 1, bias=False); self.b1 = nn.BatchNorm2d(width)
        self.c2 = nn.Conv2d(width, width, 3, stride, 1, bias=False); self.b2 = nn.BatchNorm2d(width)
        self.c3 = nn.Conv2d(width, width * 4, 1, bias=False); self.b3 = nn.BatchNorm2d(width * 4)
        self.down = down
        self.act = nn.ReLU(inplace=True)

    def forward(self, x):
        idt = x if self.down is None else self.down(x)
        y = self.act(self.b1(self.c1(x)))
        y = self.act(self.b2(self.c2(y)))
        y = self.b3(self.c3(y))
        return self.act(y + idt)


class ResNet(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), num_classes=1000):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2, 1))
        cin, stages = 64, []
        for i, n in enumerate(layers):
            width, stride = 64 << i, (1 if i == 0 else 2)
            blocks = []
            for j in range(n):
                s = stride if j == 0 else 1
                down = None
                if s != 1 or cin != width * 4:
                    down = nn.Sequential(nn.Conv2d(cin, width * 4, 1, s, bias=False), nn.BatchNorm2d(width * 4))
                blocks.append(Bottleneck(cin, width, s, down))
                cin = width * 4
            stages.append(nn.Sequential(*blocks))
        self.stages = nn.Sequential(*stages)
        self.head = nn.Linear(cin, num_classes)

    def forward(self, x):
        x = self.stages(self.stem(x))
        return self.head(torch.flatten(nn.functional.adaptive_avg_pool2d(x, 1), 1))


def resnet50(num_classes=1000):
    return ResNet((3, 4, 6, 3), num_classes)
