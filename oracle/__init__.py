"""Test-only CPU oracle of the reference's hot path.  See yolo_oracle.py."""
