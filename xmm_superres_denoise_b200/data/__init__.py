"""GPU data feed: what the reference's dataset does per sample on CPU workers after FITS decoding
(data/dataset.py:24-49,258-270; data/tools.py:79-126), per batch on the device."""
from .feed import CountsFeed, load_and_combine_simulations, reshape_img_to_res  # noqa: F401
