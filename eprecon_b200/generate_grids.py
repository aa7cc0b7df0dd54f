"""generate_grid — drop-in for ops/generate_grids.py:3-10 (strided voxel grid, x-outer / z-inner)."""
import torch


def generate_grid(n_vox, interval, device="cuda"):
    with torch.no_grad():
        axes = [torch.arange(0, n_vox[a], interval, device=device) for a in range(3)]
        grid = torch.stack(torch.meshgrid(axes[0], axes[1], axes[2], indexing="ij")).float().view(3, -1)
    return grid, (len(axes[0]), len(axes[1]), len(axes[2]))
