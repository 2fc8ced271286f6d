"""CPU oracle for the ocrs-models training hot paths. Test infrastructure only (see functional.py)."""
