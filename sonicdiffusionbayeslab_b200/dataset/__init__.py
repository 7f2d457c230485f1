from .dataset import ImageDatasetWithPrompts, SyntheticPromptDataset, synthetic_prompts  # noqa: F401
