"""Device-resident restatement of the three compression methods of the reference's `Trainer`
(neural_network_compression/common/trainer.py): `_prune_parameters` :177-193, `_reset_pruned_parameters` :195-206 and
`quantize` :42-72, for torch modules whose weights live on the GPU.

The reference pulls every tensor to the host (`layer.get_weights()`), prunes / quantizes it with NumPy + scikit-learn
and pushes it back (`layer.set_weights()`), every batch.  Here the parameters never leave the device: the same helper
calls (`utility.prune_weigth`, `utility.get_weight_distribution`, `utility.get_quantized_weight`) run in place on the
parameter storage.  Training loops, optimizers, data and reports are out of scope.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

from . import utility


def _layer_params(layer: torch.nn.Module):
    """(kernel, bias) of a layer, like Keras `layer.get_weights()` -- `[]` for layers without parameters."""
    w = getattr(layer, "weight", None)
    b = getattr(layer, "bias", None)
    return [p for p in (w, b) if p is not None]


class Compressor:
    """The mask cache and the prune / re-apply / quantize drivers of `Trainer`, over a `torch.nn.Module`.

    layers_to_prune_with_threshold: {layer: (weight_threshold, bias_threshold)} -- the reference's
    `_layers_to_prune_with_threshold` (le_net_300_100_trainer.py:22-27).
    """

    def __init__(self, neural_network: torch.nn.Module,
                 layers_to_prune_with_threshold: Dict[torch.nn.Module, Tuple[float, float]]):
        self.neural_network = neural_network
        self._layers_to_prune_with_threshold = layers_to_prune_with_threshold
        # trainer.py:25 (a class attribute there; per instance here)
        self.pruned_indexes_by_layer: Dict[torch.nn.Module, Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]] = {}

    @torch.no_grad()
    def _prune_parameters(self, with_standard_deviation_smoothing: bool) -> None:
        """trainer.py:177-193: prune kernel and bias of every configured layer in place, remember the masks."""
        for layer, (weight_threshold, bias_threshold) in self._layers_to_prune_with_threshold.items():
            zero_weight_indexes = utility.prune_weigth(layer.weight.data, threshold=weight_threshold,
                                                       std_smooth=with_standard_deviation_smoothing)
            zero_bias_indexes = None
            if layer.bias is not None:
                zero_bias_indexes = utility.prune_weigth(layer.bias.data, threshold=bias_threshold,
                                                         std_smooth=with_standard_deviation_smoothing)
            self.pruned_indexes_by_layer[layer] = (zero_weight_indexes, zero_bias_indexes)

    @torch.no_grad()
    def _reset_pruned_parameters(self, gradients_too: bool = False) -> None:
        """trainer.py:195-206: weights[mask] = 0 after the optimizer step (optionally the gradients as well, which
        keeps optimizer moments of pruned weights at zero -- the 'masked gradient apply' of the north star)."""
        for layer, (zero_weight_indexes, zero_bias_indexes) in self.pruned_indexes_by_layer.items():
            utility.apply_mask(layer.weight.data, zero_weight_indexes)
            if gradients_too and layer.weight.grad is not None:
                utility.apply_mask(layer.weight.grad, zero_weight_indexes)
            if layer.bias is not None and zero_bias_indexes is not None:
                utility.apply_mask(layer.bias.data, zero_bias_indexes)
                if gradients_too and layer.bias.grad is not None:
                    utility.apply_mask(layer.bias.grad, zero_bias_indexes)

    @torch.no_grad()
    def quantize(self, with_cumulative_weight_distribution: bool, maximum_centroid_bits: int,
                 k_means_initialization_mode: str, layers: Optional[Iterable[torch.nn.Module]] = None):
        """trainer.py:42-72: k-means weight sharing of every kernel and bias of every layer of the model's config.
        Returns {layer: [KMeansResult or None per parameter]} (the reference returns the test accuracy, which needs
        the data pipeline that is out of scope)."""
        if layers is None:
            cfg = getattr(self.neural_network, "get_config", None)
            layers = list(cfg().values()) if cfg is not None else [m for m in self.neural_network.modules() if _layer_params(m) and not list(m.children())]
        fitted = {}
        for layer in layers:
            models = []
            for params in _layer_params(layer):
                cdfs = None
                if with_cumulative_weight_distribution:
                    # trainer.py:55-60: CDF of the non-zero weights (fused survivor selection, nothing materialised)
                    cdfs = utility.get_weight_distribution(params.data, skip_zeros=True)
                quantized, kmeans = utility.get_quantized_weight(params.data, bits=maximum_centroid_bits,
                                                                 mode=k_means_initialization_mode, cdfs=cdfs)
                if kmeans is not None:
                    params.data.copy_(quantized)  # layer.set_weights(quantized) (trainer.py:70)
                models.append(kmeans)
            fitted[layer] = models
        return fitted

    @torch.no_grad()
    def trained_quantization_step(self, fitted, learning_rate: float) -> None:
        """The fine-tuning step of Deep Compression's trained quantization, which the reference describes but leaves
        unimplemented (papers/lat/report.tex:149-158): the gradient of a shared weight is the sum of the gradients of
        the weights that carry its index, dL/dC_k = sum_ij dL/dW_ij 1(I_ij = k) (report.tex:152); the codebook takes a
        plain SGD step and the layer is re-materialised from the unchanged packed indices.

        fitted: what quantize() returned; every parameter must hold its loss gradient in .grad."""
        for layer, models in fitted.items():
            for params, kmeans in zip(_layer_params(layer), models):
                if kmeans is None or params.grad is None:
                    continue
                k, bits = kmeans.n_clusters, kmeans.code_bits
                grad = params.grad.contiguous().view(-1)
                g = utility.cluster_gradient_sum(grad, kmeans.packed_codes, k, bits)  # float64[k], exact per tile
                centers = (kmeans.cluster_centers_.astype("float64").ravel() - learning_rate * g).astype("float32")
                kmeans.cluster_centers_ = centers.reshape(-1, 1)
                params.data.copy_(utility.dequantize(kmeans.packed_codes, params.numel(), bits, centers).view_as(params))

