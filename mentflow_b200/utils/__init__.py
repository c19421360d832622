"""Small host-side helpers with the reference's names (mentflow/utils/{grid,utils}.py)."""
import itertools
import pickle

import torch


def coords_from_edges(edges: torch.Tensor) -> torch.Tensor:
    """Bin centres (utils/grid.py:5-6)."""
    return 0.5 * (edges[:-1] + edges[1:])


def get_grid_points(*coords: torch.Tensor) -> torch.Tensor:
    """(prod(len), ndim) points of a regular grid, 'ij' order (utils/grid.py:9-10)."""
    mesh = torch.meshgrid(*coords, indexing="ij")
    return torch.stack([m.reshape(-1) for m in mesh], dim=1)


def unravel(iterable):
    """Flatten one nesting level (utils/utils.py:10-11)."""
    return itertools.chain.from_iterable(iterable)


def grab(x: torch.Tensor):
    """Tensor -> numpy on the host (utils/utils.py:14-15)."""
    return x.detach().cpu().numpy()


def load_pickle(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def save_pickle(obj, path):
    with open(path, "wb") as f:
        pickle.dump(obj, f)
