"""Import shim: the reference's package name. Checkpoints written by HoagyC/sparse_coding pickle
``autoencoders.learned_dict.TiedSAE`` etc.; callers such as big_sweep.py / basic_l1_sweep.py do
``from autoencoders.sae_ensemble import FunctionalTiedSAE``. These modules re-export the engine-backed classes
from ``sparse_coding_b200`` under those names."""
