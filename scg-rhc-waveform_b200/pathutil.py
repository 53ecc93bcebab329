"""The data-root seam of the window-preparation path (reference: pathutil.py:4-6).  ``SCG_RHC_DATA_PATH``
overrides the author's absolute path.  The reference's directory-clearing helpers (pathutil.py:9-19) are
outside the path and are not provided."""
import os

DATA_PATH = os.environ.get('SCG_RHC_DATA_PATH', os.path.join('/', 'home', 'jesse', 'scg-rhc-database'))

PROCESSED_DATA_PATH = os.path.join(DATA_PATH, 'processed_data')
