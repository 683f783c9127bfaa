from .efficient_ensemble_merged import EfficientEnsembleMerged  # noqa: F401
