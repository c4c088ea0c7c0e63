from .data_processor import DataProcessor


class BachDataProcessor(DataProcessor):
    """No override (VQCPCB/data_processor/bach_data_processor.py:7-12)."""
