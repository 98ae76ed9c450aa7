"""Minimal FASTA reader with the interface of the reference's test helper
(src/pyfastani/_fasta.pyx:41-103): ``Parser(path)`` iterates ``Record(id, seq)``."""
import gzip


class Record:
    __slots__ = ("id", "seq")

    def __init__(self, id, seq):
        self.id = id
        self.seq = seq


class Parser:
    def __init__(self, path):
        self.path = path

    def __iter__(self):
        opener = gzip.open if str(self.path).endswith(".gz") else open
        name, chunks = None, []
        with opener(self.path, "rt") as handle:
            for line in handle:
                line = line.strip()
                if line.startswith(">"):
                    if name is not None:
                        yield Record(name, "".join(chunks).upper())
                    name, chunks = line[1:].split()[0] if len(line) > 1 else "", []
                elif line:
                    chunks.append(line)
        if name is not None:
            yield Record(name, "".join(chunks).upper())
