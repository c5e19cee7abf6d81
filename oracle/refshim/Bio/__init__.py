"""TEST INFRASTRUCTURE ONLY -- empty stand-in for biopython."""
Seq = SeqIO = SeqRecord = None
