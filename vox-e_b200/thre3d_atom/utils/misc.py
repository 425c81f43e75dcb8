"""Small host-side helpers (reference: thre3d_atom/utils/misc.py:10-58).  ``log_config_to_disk`` is kept because every
training / editing script of the reference imports it from here at top level; it takes any mapping (the scripts pass an
``EasyDict``), so this module does not import ``easydict``."""
from pathlib import Path
from typing import Any, Callable, List, Mapping, Optional, Sequence, Tuple

import numpy as np


def check_power_of_2(x: int) -> bool:
    return x & (x - 1) == 0


def batchify(
    processor_fn: Callable[..., Any],
    collate_fn: Callable[[Sequence[Any]], Any],
    chunk_size: Optional[int] = None,
    verbose: bool = False,
) -> Callable[..., Any]:
    """Wrap ``processor_fn`` so that its first argument is processed in slices of ``chunk_size`` and the partial results
    are merged with ``collate_fn``; ``chunk_size=None`` returns the function untouched."""
    if chunk_size is None:
        return processor_fn

    def chunked(inputs: Sequence[Any], *args, **kwargs) -> Any:
        starts = range(0, len(inputs), chunk_size)
        if verbose:
            from tqdm import tqdm

            starts = tqdm(starts)
        return collate_fn([processor_fn(inputs[s : s + chunk_size], *args, **kwargs) for s in starts])

    return chunked


def compute_thre3d_grid_sizes(
    final_required_resolution: Tuple[int, int, int], num_stages: int, scale_factor: float
) -> List[Tuple[int, int, int]]:
    """Coarse-to-fine grid sizes for progressive training, finest last."""
    sizes = [tuple(int(v) for v in final_required_resolution)]
    for _ in range(num_stages - 1):
        sizes.insert(0, tuple(int(np.ceil(v / scale_factor)) for v in sizes[0]))
    return sizes


def log_config_to_disk(args: Mapping[str, Any], output_dir: Path, config_file_name: str = "config.yml") -> None:
    """Write the run configuration as block-style YAML into ``output_dir`` (created if missing)."""
    import yaml

    output_dir = Path(output_dir)
    output_dir.mkdir(exist_ok=True, parents=True)
    with open(output_dir / config_file_name, "w") as handle:
        yaml.dump(dict(args), handle, default_flow_style=False)
