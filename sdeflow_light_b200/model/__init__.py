"""2-D U-Net building blocks (drop-in for the reference's model/ package)."""
