from ._evolution import Population, rescale, to_phenotype
