"""Predictor training-data path."""
