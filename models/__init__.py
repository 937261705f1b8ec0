"""Drop-in import path of the reference (`from models.rendering import render`, `from models.networks import NGP`,
`from models.custom_functions import ...`): thin re-exports of ar_nerf_b200."""
