"""Glue that lets the reference's own scripts run unmodified on this package."""
