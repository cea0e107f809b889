"""``pbnn.utils`` mirror for the pieces that sit right after the sampler on the hot path."""
from .misc import save_results, thinning_fn  # noqa: F401
