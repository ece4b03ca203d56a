"""B200-native log-mel front end + cosine scoring (drop-in for the hot path of
yuriyvnv/speech_transcript_embeddings).  See DESIGN.md."""
__version__ = "0.1.0"
