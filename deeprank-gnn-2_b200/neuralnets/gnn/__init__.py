"""GNN architectures with the reference's class names, constructor signatures and state_dict keys."""
