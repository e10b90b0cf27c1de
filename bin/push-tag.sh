#!/usr/bin/env bash
# Tag the current commit with the package version and push the tag.
set -e
version=$(python -c "import re; print(re.search(r'version = \"(.+?)\"', open('pyproject.toml').read()).group(1))")
git diff-index --quiet HEAD || { echo "working tree not clean"; exit 1; }
git tag "v$version"
git push origin "v$version"
