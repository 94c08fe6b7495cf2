"""Four-symbol stand-in for MONAI so the *unmodified* reference model files import.

TEST INFRASTRUCTURE ONLY (see oracle/README.md). The reference
(`/root/reference/medimgen/*_with_strides.py:40-42`) needs exactly
`monai.networks.blocks.{Convolution, MLPBlock}`, `monai.networks.layers.factories.Pool`
and `monai.utils.ensure_tuple_rep`. MONAI itself is not installable here (no network),
so these restate the published behaviour of those four symbols.
"""
