"""Oracle = test infrastructure. CPU restatement of the reference's hot path; never imported by the product package."""
