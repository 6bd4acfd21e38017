class optModel:
    """placeholder base so isinstance checks pass"""
    modelSense = None
