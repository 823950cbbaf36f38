"""Test infrastructure: CPU oracle of the K-FAC hot path. Never imported by the product package."""
