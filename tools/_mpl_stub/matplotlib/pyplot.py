"""see __init__.py — plotting is never reached on the hot path."""
