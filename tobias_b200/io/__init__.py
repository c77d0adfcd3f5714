"""Host-side file formats on either side of the hot path (FASTA, BAM/BAI, bigWig, BED)."""
