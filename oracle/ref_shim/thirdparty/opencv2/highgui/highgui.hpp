// OpenCV is absent from the image; the reference's layer sources include it but the compiled paths use nothing of it. TEST INFRASTRUCTURE for oracle/_ref.
