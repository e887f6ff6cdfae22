"""Drop-in for the reference's ``models.layers.SAST`` package, backed by sast_b200."""
