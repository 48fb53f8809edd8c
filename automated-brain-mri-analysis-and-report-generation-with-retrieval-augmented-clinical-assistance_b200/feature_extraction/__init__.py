"""Drop-ins for the voxel statistics of the reference's feature_extraction/ (utils, step3, step4)."""
