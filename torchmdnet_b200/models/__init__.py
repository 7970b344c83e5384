__all_models__ = ["tensornet", "equivariant-transformer"]
