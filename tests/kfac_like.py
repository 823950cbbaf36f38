"""Generator of K-FAC-like factor sequences (test helper): F_{t+1} = 0.95 F_t + 0.05 X_t^T X_t / m with
F_0 = I and fresh samples X_t of m < n rows (ReLU-like mixed features with graded scales and a bias column):
a decaying identity cluster plus low-rank updates -- the spectrum class the bench's ResNet-50 factors have."""
import torch


def kfac_like_sequence(n, m, steps, device, seed=7):
    g = torch.Generator(device='cpu').manual_seed(seed)
    scale = torch.logspace(0, -2, n).unsqueeze(0)
    mix = torch.randn(n, n, generator=g) / n ** 0.5
    F = torch.eye(n, device=device)
    for _ in range(steps):
        x = torch.relu(torch.randn(m, n, generator=g) @ mix + 0.3) * scale
        x[:, -1] = 1.0
        x = x.to(device)
        F = 0.95 * F + 0.05 * (x.t() @ x) / m
        F = ((F + F.t()) / 2).contiguous()
        yield F
