"""Minimal stand-in for `torchmetrics` (absent from this image, no network): only what the reference's
trainer imports, `torchmetrics.image.StructuralSimilarityIndexMeasure`."""
