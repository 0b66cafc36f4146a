"""Mirror of the reference package layout so that `from clip.loss import ClipLoss` call sites only change the root."""
