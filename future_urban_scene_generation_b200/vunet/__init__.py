"""Drop-in mirror of the reference's `vunet` package for the hot path."""
